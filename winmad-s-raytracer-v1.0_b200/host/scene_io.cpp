// Scene ingestion, camera set-up and image output on the host.
//
// Restates, with its own small XML and OBJ readers (the reference vendors TinyXML and tinyobjloader):
//   Scene::loadScene(char*)      R/src/scene/scene.cpp:259-467   element order, positional children,
//                                                                 height->xResolution / width->yResolution
//   tinyobj::LoadObj             R/src/tinyobjloader/tiny_obj_loader.cpp:460-661 ('v', 'f' with fan
//                                triangulation :227-239, 'g'/'o' shape breaks, (float)atof parsing :91-97)
//   Camera::setup / generateRay  R/src/scene/camera.cpp:3-42, Transform::tPoint transform.h:126-140
//   ImageFilm::outputImage       R/src/scene/film.cpp:39-64 (scale, clamp, gamma 1/2.2, 8-bit)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <climits>
#include <fstream>
#include <sstream>
#include "host_scene.h"

namespace wrt {

// ------------------------------------------------------------------------------------------------
// minimal XML: elements + attributes only (all the .scene format uses)
// ------------------------------------------------------------------------------------------------
namespace {

struct XmlNode {
    std::string name;
    std::vector<std::pair<std::string, std::string> > attrs;
    std::vector<XmlNode> kids;
    const std::string* attr(const char* key) const {
        for (size_t i = 0; i < attrs.size(); i++) if (attrs[i].first == key) return &attrs[i].second;
        return 0;
    }
    // TiXmlElement::Attribute(name, double*): value via atof; untouched when missing
    void get(const char* key, double* v) const { const std::string* s = attr(key); if (s) *v = atof(s->c_str()); }
    void get(const char* key, int* v) const { const std::string* s = attr(key); if (s) *v = atoi(s->c_str()); }
};

struct XmlParser {
    const std::string& s;
    size_t i;
    explicit XmlParser(const std::string& src) : s(src), i(0) {}
    void skip_ws() { while (i < s.size() && isspace((unsigned char)s[i])) i++; }
    bool starts(const char* p) const { return s.compare(i, strlen(p), p) == 0; }
    void skip_misc() {  // whitespace, comments, declarations, text
        for (;;) {
            while (i < s.size() && s[i] != '<') i++;
            if (i >= s.size()) return;
            if (starts("<!--")) { size_t e = s.find("-->", i + 4); i = (e == std::string::npos) ? s.size() : e + 3; continue; }
            if (starts("<?")) { size_t e = s.find("?>", i + 2); i = (e == std::string::npos) ? s.size() : e + 2; continue; }
            if (starts("<!")) { size_t e = s.find('>', i + 2); i = (e == std::string::npos) ? s.size() : e + 1; continue; }
            return;
        }
    }
    bool parse_element(XmlNode& out) {
        skip_misc();
        if (i >= s.size() || s[i] != '<' || starts("</")) return false;
        i++;
        size_t b = i;
        while (i < s.size() && !isspace((unsigned char)s[i]) && s[i] != '>' && s[i] != '/') i++;
        out.name = s.substr(b, i - b);
        for (;;) {
            skip_ws();
            if (i >= s.size()) return true;
            if (s[i] == '/') { i = s.find('>', i); i = (i == std::string::npos) ? s.size() : i + 1; return true; }
            if (s[i] == '>') { i++; break; }
            size_t kb = i;
            while (i < s.size() && s[i] != '=' && !isspace((unsigned char)s[i]) && s[i] != '>' && s[i] != '/') i++;
            std::string key = s.substr(kb, i - kb);
            skip_ws();
            std::string val;
            if (i < s.size() && s[i] == '=') {
                i++; skip_ws();
                if (i < s.size() && (s[i] == '"' || s[i] == '\'')) {
                    char q = s[i++];
                    size_t vb = i;
                    while (i < s.size() && s[i] != q) i++;
                    val = s.substr(vb, i - vb);
                    if (i < s.size()) i++;
                }
            }
            out.attrs.push_back(std::make_pair(key, val));
        }
        for (;;) {  // children until the closing tag
            skip_misc();
            if (i >= s.size()) return true;
            if (starts("</")) { i = s.find('>', i); i = (i == std::string::npos) ? s.size() : i + 1; return true; }
            XmlNode kid;
            if (!parse_element(kid)) return true;
            out.kids.push_back(kid);
        }
    }
};

bool read_file(const char* path, std::string& out)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::stringstream ss; ss << f.rdbuf();
    out = ss.str();
    return true;
}

std::string base_name(const std::string& p)
{
    size_t k = p.find_last_of("/\\");
    return (k == std::string::npos) ? p : p.substr(k + 1);
}

std::string dir_name(const std::string& p)
{
    size_t k = p.find_last_of('/');
    return (k == std::string::npos) ? std::string(".") : p.substr(0, k);
}

// The reference opens the path exactly as written and silently adds no geometry when that fails
// (scene.cpp:337, LoadObj's error string is dropped).  torus.scene ships with absolute Windows
// paths, so after the literal path we also try the file's base name under $WRT_OBJ_DIR, next to
// the scene file, and in <scene dir>/ObjFiles.  A file found nowhere contributes nothing — as in
// the reference.
std::string resolve_obj(const std::string& given, const std::string& scene_dir)
{
    std::ifstream f(given.c_str());
    if (f) return given;
    std::string b = base_name(given);
    std::vector<std::string> cand;
    if (const char* e = getenv("WRT_OBJ_DIR")) cand.push_back(std::string(e) + "/" + b);
    cand.push_back(scene_dir + "/" + b);
    cand.push_back(scene_dir + "/ObjFiles/" + b);
    for (size_t i = 0; i < cand.size(); i++) { std::ifstream g(cand[i].c_str()); if (g) return cand[i]; }
    return given;
}

inline float parse_float(const char*& tok)  // tiny_obj_loader.cpp:91-97: double parse, then narrow
{
    tok += strspn(tok, " \t");
    float f = (float)atof(tok);
    tok += strcspn(tok, " \t\r");
    return f;
}

inline int fix_index(int idx, int n) { return idx > 0 ? idx - 1 : (idx == 0 ? 0 : n + idx); }  // :66-78

void read_vec3(const XmlNode& e, float v[3])
{
    double x = 0, y = 0, z = 0;
    e.get("x", &x); e.get("y", &y); e.get("z", &z);
    v[0] = (float)x; v[1] = (float)y; v[2] = (float)z;
}

void read_rgb(const XmlNode& e, float v[3])
{
    double r = 0, g = 0, b = 0;
    e.get("r", &r); e.get("g", &g); e.get("b", &b);
    v[0] = (float)r; v[1] = (float)g; v[2] = (float)b;
}

void add_prim(HostScene& hs, int kind, const float* d9, int matid)
{
    hs.prim_kind.push_back(kind);
    hs.prim_data.insert(hs.prim_data.end(), d9, d9 + 9);
    hs.prim_matid.push_back(matid);
}

}  // namespace

// OBJ reader: positions and faces only (that is all Scene::loadScene consumes).  One float list of
// triangles (9 floats each) per shape, in tinyobj's shape/face/fan order.
bool load_obj_triangles(const char* path, std::vector<std::vector<float> >& shapes,
                        std::vector<std::string>& names)
{
    shapes.clear(); names.clear();
    std::ifstream ifs(path);
    if (!ifs) return false;
    std::vector<float> v;
    std::vector<float> cur;
    std::string name;
    std::string line;
    bool any_face = false;
    while (std::getline(ifs, line)) {
        if (!line.empty() && line[line.size() - 1] == '\n') line.erase(line.size() - 1);
        if (line.empty()) continue;
        const char* tok = line.c_str();
        tok += strspn(tok, " \t");
        if (tok[0] == '\0' || tok[0] == '#') continue;
        if (tok[0] == 'v' && (tok[1] == ' ' || tok[1] == '\t')) {
            tok += 2;
            float x = parse_float(tok), y = parse_float(tok), z = parse_float(tok);
            v.push_back(x); v.push_back(y); v.push_back(z);
            continue;
        }
        if (tok[0] == 'f' && (tok[1] == ' ' || tok[1] == '\t')) {
            tok += 2;
            tok += strspn(tok, " \t");
            std::vector<int> face;
            while (!(tok[0] == '\r' || tok[0] == '\n' || tok[0] == '\0')) {
                face.push_back(fix_index(atoi(tok), (int)(v.size() / 3)));
                tok += strcspn(tok, " \t\r");   // skip the rest of the i/j/k triple
                tok += strspn(tok, " \t\r");
            }
            any_face = true;
            for (size_t k = 2; k < face.size(); k++) {  // fan: (f0, f[k-1], f[k])
                int idx[3] = { face[0], face[k - 1], face[k] };
                for (int c = 0; c < 3; c++) {
                    size_t o = 3 * (size_t)idx[c];
                    if (o + 2 < v.size()) { cur.push_back(v[o]); cur.push_back(v[o + 1]); cur.push_back(v[o + 2]); }
                    else { cur.push_back(0.f); cur.push_back(0.f); cur.push_back(0.f); }
                }
            }
            continue;
        }
        const bool is_g = tok[0] == 'g' && (tok[1] == ' ' || tok[1] == '\t');
        const bool is_o = tok[0] == 'o' && (tok[1] == ' ' || tok[1] == '\t');
        if (is_g || is_o) {  // flush the previous face group as a shape (:609-655)
            if (any_face) { shapes.push_back(cur); names.push_back(name); }
            cur.clear(); any_face = false;
            char buf[4096]; buf[0] = 0;
            sscanf(tok + 2, "%4095s", buf);
            name = buf;
            continue;
        }
        // mtllib with a missing file aborts the reference's load (:593-603) — none of the scenes
        // this path targets use it; material data is never consumed by Scene::loadScene.
    }
    if (any_face) { shapes.push_back(cur); names.push_back(name); }
    return true;
}

bool load_scene_file(const char* path, HostScene& hs, std::string& err)
{
    std::string text;
    if (!read_file(path, text)) { err = std::string("cannot open scene file ") + path; return false; }
    XmlParser xp(text);
    XmlNode root;
    if (!xp.parse_element(root)) { err = "scene file has no root element"; return false; }
    const std::string sdir = dir_name(path);

    for (size_t k = 0; k < root.kids.size(); k++) {
        const XmlNode& it = root.kids[k];
        const std::vector<XmlNode>& c = it.kids;
        if (it.name == "camera" && c.size() >= 5) {  // scene.cpp:276-304, children are positional
            float pos[3], fwd[3], up[3];
            read_vec3(c[0], pos); read_vec3(c[1], fwd); read_vec3(c[2], up);
            double x = 0, y = 0, fov = 0;
            c[3].get("height", &x); c[3].get("width", &y);  // height -> xResolution, width -> yResolution
            c[4].get("horizontalFOV", &fov);
            for (int a = 0; a < 3; a++) { hs.cam_args[a] = pos[a]; hs.cam_args[3 + a] = fwd[a]; hs.cam_args[6 + a] = up[a]; }
            hs.cam_args[9] = (float)x; hs.cam_args[10] = (float)y; hs.cam_args[11] = (float)fov;
            camera_setup(pos, fwd, up, (float)x, (float)y, (float)fov, &hs.camera);
            hs.has_camera = true;
        } else if (it.name == "material" && c.size() >= 5) {  // :305-332
            float m[11];
            read_rgb(c[0], m + 0); read_rgb(c[1], m + 3); read_rgb(c[2], m + 7);
            double pe = 1, idx = -1;
            c[3].get("phongExp", &pe); c[4].get("refracIndex", &idx);
            m[6] = (float)pe; m[10] = (float)idx;
            hs.materials.insert(hs.materials.end(), m, m + 11);
        } else if (it.name == "object" && c.size() >= 2) {  // :333-375
            const std::string* p = c[0].attr("path");
            int id = 0; c[1].get("matid", &id);
            std::vector<std::vector<float> > shapes; std::vector<std::string> names;
            if (p) load_obj_triangles(resolve_obj(*p, sdir).c_str(), shapes, names);
            for (size_t s = 0; s < shapes.size(); s++)
                for (size_t f = 0; f + 8 < shapes[s].size(); f += 9) {
                    float t[9]; memcpy(t, &shapes[s][f], sizeof t);
                    if (names[s] == "water") {  // :359-368
                        float e1[3] = { t[3] - t[0], t[4] - t[1], t[5] - t[2] };
                        float e2[3] = { t[6] - t[0], t[7] - t[1], t[8] - t[2] };
                        float ny = e1[2] * e2[0] - e1[0] * e2[2];
                        if (ny < kEps) for (int a = 0; a < 3; a++) std::swap(t[a], t[6 + a]);
                    }
                    add_prim(hs, WRT_PRIM_TRIANGLE, t, id);
                }
        } else if (it.name == "sphere" && c.size() >= 3) {  // :376-396
            float o[3]; read_vec3(c[0], o);
            double radius = 0; c[1].get("radius", &radius);
            int id = 0; c[2].get("matid", &id);
            float d[9] = { o[0], o[1], o[2], (float)radius, 0, 0, 0, 0, 0 };
            add_prim(hs, WRT_PRIM_SPHERE, d, id);
        } else if (it.name == "area_light" && c.size() >= 2) {  // :397-432
            const std::string* p = c[0].attr("path");
            float inten[3]; read_rgb(c[1], inten);
            std::vector<std::vector<float> > shapes; std::vector<std::string> names;
            if (p) load_obj_triangles(resolve_obj(*p, sdir).c_str(), shapes, names);
            for (size_t s = 0; s < shapes.size(); s++)
                for (size_t f = 0; f + 8 < shapes[s].size(); f += 9) {
                    const float* t = &shapes[s][f];
                    float l[12]; memcpy(l, t, 9 * sizeof(float)); l[9] = inten[0]; l[10] = inten[1]; l[11] = inten[2];
                    hs.lights.insert(hs.lights.end(), l, l + 12);
                    add_prim(hs, WRT_PRIM_TRIANGLE, t, -((int)(f / 9) + 1));  // matId = -(f+1), :427
                }
        }
        // <homo_media>: participating media are outside this path.
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// camera
// ------------------------------------------------------------------------------------------------
namespace {

struct M4 { float m[4][4]; };

M4 ident() { M4 r; memset(&r, 0, sizeof r); r.m[0][0] = r.m[1][1] = r.m[2][2] = r.m[3][3] = 1.f; return r; }

M4 mul(const M4& a, const M4& b)  // operator*, transform.cpp:27-37
{
    M4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] + a.m[i][3] * b.m[3][j];
    return r;
}

// inverse(Matrix4x4), transform.cpp:48-173: adjugate / determinant with every cofactor written as six signed triple
// products summed left to right.  Camera matrices must be BIT-identical to the reference's (primary rays of scenes with
// coordinates in the thousands move by whole ulps otherwise), so the term order of the reference is kept as a table:
// inv[k] = sum over t of sign * (m[a] * m[b]) * m[c], k and a, b, c indexing the row-major 16 floats.
const signed char kCofactor[16][6][4] = {
    { { 1, 5,10,15}, {-1, 5,11,14}, {-1, 9, 6,15}, { 1, 9, 7,14}, { 1,13, 6,11}, {-1,13, 7,10} },
    { {-1, 1,10,15}, { 1, 1,11,14}, { 1, 9, 2,15}, {-1, 9, 3,14}, {-1,13, 2,11}, { 1,13, 3,10} },
    { { 1, 1, 6,15}, {-1, 1, 7,14}, {-1, 5, 2,15}, { 1, 5, 3,14}, { 1,13, 2, 7}, {-1,13, 3, 6} },
    { {-1, 1, 6,11}, { 1, 1, 7,10}, { 1, 5, 2,11}, {-1, 5, 3,10}, {-1, 9, 2, 7}, { 1, 9, 3, 6} },
    { {-1, 4,10,15}, { 1, 4,11,14}, { 1, 8, 6,15}, {-1, 8, 7,14}, {-1,12, 6,11}, { 1,12, 7,10} },
    { { 1, 0,10,15}, {-1, 0,11,14}, {-1, 8, 2,15}, { 1, 8, 3,14}, { 1,12, 2,11}, {-1,12, 3,10} },
    { {-1, 0, 6,15}, { 1, 0, 7,14}, { 1, 4, 2,15}, {-1, 4, 3,14}, {-1,12, 2, 7}, { 1,12, 3, 6} },
    { { 1, 0, 6,11}, {-1, 0, 7,10}, {-1, 4, 2,11}, { 1, 4, 3,10}, { 1, 8, 2, 7}, {-1, 8, 3, 6} },
    { { 1, 4, 9,15}, {-1, 4,11,13}, {-1, 8, 5,15}, { 1, 8, 7,13}, { 1,12, 5,11}, {-1,12, 7, 9} },
    { {-1, 0, 9,15}, { 1, 0,11,13}, { 1, 8, 1,15}, {-1, 8, 3,13}, {-1,12, 1,11}, { 1,12, 3, 9} },
    { { 1, 0, 5,15}, {-1, 0, 7,13}, {-1, 4, 1,15}, { 1, 4, 3,13}, { 1,12, 1, 7}, {-1,12, 3, 5} },
    { {-1, 0, 5,11}, { 1, 0, 7, 9}, { 1, 4, 1,11}, {-1, 4, 3, 9}, {-1, 8, 1, 7}, { 1, 8, 3, 5} },
    { {-1, 4, 9,14}, { 1, 4,10,13}, { 1, 8, 5,14}, {-1, 8, 6,13}, {-1,12, 5,10}, { 1,12, 6, 9} },
    { { 1, 0, 9,14}, {-1, 0,10,13}, {-1, 8, 1,14}, { 1, 8, 2,13}, { 1,12, 1,10}, {-1,12, 2, 9} },
    { {-1, 0, 5,14}, { 1, 0, 6,13}, { 1, 4, 1,14}, {-1, 4, 2,13}, {-1,12, 1, 6}, { 1,12, 2, 5} },
    { { 1, 0, 5,10}, {-1, 0, 6, 9}, {-1, 4, 1,10}, { 1, 4, 2, 9}, { 1, 8, 1, 6}, {-1, 8, 2, 5} },
};

M4 inverse(const M4& a)
{
    const float* m = &a.m[0][0];
    float inv[16];
    for (int k = 0; k < 16; k++) {
        float acc = 0.f;
        for (int t = 0; t < 6; t++) {
            const signed char* q = kCofactor[k][t];
            const float prod = m[q[1]] * m[q[2]] * m[q[3]];              // left to right, as written in the reference
            if (t == 0) acc = q[0] < 0 ? -prod : prod;                    // (-x) * y * z == -(x * y * z) exactly
            else acc = q[0] < 0 ? acc - prod : acc + prod;
        }
        inv[k] = acc;
    }
    float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    det = 1.f / det;
    M4 r;
    float* o = &r.m[0][0];
    for (int k = 0; k < 16; k++) o[k] = inv[k] * det;
    return r;
}

// Transform (transform.h:69-123): a matrix and its inverse carried side by side; the product multiplies the
// inverses in the opposite order instead of inverting the product (transform.cpp:198-203).
struct T4 { M4 m, inv; };
T4 t4(const M4& m) { T4 r; r.m = m; r.inv = inverse(m); return r; }
T4 t4(const M4& m, const M4& inv) { T4 r; r.m = m; r.inv = inv; return r; }

void normalize3(float v[3])
{
    float len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] /= len; v[1] /= len; v[2] /= len;
}

void cross3(const float a[3], const float b[3], float r[3])
{
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}

float dot3(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

M4 mul(const M4& a, const M4& b);
T4 mul(const T4& a, const T4& b) { return t4(mul(a.m, b.m), mul(b.inv, a.inv)); }
T4 scale_t(float x, float y, float z)        // scale(), transform.cpp:276-288
{
    M4 m = ident(), i = ident();
    m.m[0][0] = x; m.m[1][1] = y; m.m[2][2] = z;
    i.m[0][0] = 1.0f / x; i.m[1][1] = 1.0f / y; i.m[2][2] = 1.0f / z;
    return t4(m, i);
}
T4 translate_t(float x, float y, float z)    // translate(), transform.cpp:262-274
{
    M4 m = ident(), i = ident();
    m.m[0][3] = x; m.m[1][3] = y; m.m[2][3] = z;
    i.m[0][3] = -x; i.m[1][3] = -y; i.m[2][3] = -z;
    return t4(m, i);
}

// Transform::tPoint, transform.h:126-140 (+ Vector3 operator/ returning INF on |w| <= EPS, vector.cpp:35-39)
void t_point(const float m[16], const float p[3], float out[3])
{
    float x = p[0], y = p[1], z = p[2];
    float xp = m[0] * x + m[1] * y + m[2] * z + m[3];
    float yp = m[4] * x + m[5] * y + m[6] * z + m[7];
    float zp = m[8] * x + m[9] * y + m[10] * z + m[11];
    float wp = m[12] * x + m[13] * y + m[14] * z + m[15];
    if (cmp_eps(wp - 1.0f) == 0) { out[0] = xp; out[1] = yp; out[2] = zp; }
    else if (cmp_eps(wp) == 0) { out[0] = out[1] = out[2] = kInf; }
    else { out[0] = xp / wp; out[1] = yp / wp; out[2] = zp / wp; }
}

}  // namespace

// Camera::setup, camera.cpp:3-29 (lookAt transform.cpp:353-370, perspective :379-387), operation for operation:
// the matrices are bit-identical to the reference's (tests/test_host.py::test_camera_setup_bit_identical_to_reference).
void camera_setup(const float pos_in[3], const float fwd_in[3], const float up_in[3], float xres, float yres,
                  float fov, wrt_camera* out)
{
    const float pi = std::acos(-1.0f);
    float pos[3] = { pos_in[0], pos_in[1], pos_in[2] };
    float fwd[3] = { fwd_in[0], fwd_in[1], fwd_in[2] };
    float up[3] = { up_in[0], up_in[1], up_in[2] };
    normalize3(fwd); normalize3(up);

    // lookAt(pos, pos + forward, up): _dir = look - pos (normalised), _up = up x (-_dir) (normalised), _left = _up x _dir;
    // rows (_up, -_up.pos), (_left, -_left.pos), (-_dir, +_dir.pos... as -(-_dir.pos))
    float look[3] = { pos[0] + fwd[0], pos[1] + fwd[1], pos[2] + fwd[2] };
    float dir[3] = { look[0] - pos[0], look[1] - pos[1], look[2] - pos[2] };
    normalize3(dir);
    float ndir[3] = { -dir[0], -dir[1], -dir[2] };
    float ux[3]; cross3(up, ndir, ux); normalize3(ux);
    float lf[3]; cross3(ux, dir, lf);
    M4 w2c_m = ident();
    const float px = dot3(ux, pos), py = dot3(lf, pos), pz = dot3(ndir, pos);
    for (int a = 0; a < 3; a++) { w2c_m.m[0][a] = ux[a]; w2c_m.m[1][a] = lf[a]; w2c_m.m[2][a] = ndir[a]; }
    w2c_m.m[0][3] = -px; w2c_m.m[1][3] = -py; w2c_m.m[2][3] = -pz;
    const T4 w2c = t4(w2c_m);

    // perspective(fov, 0.1, 10000) = scale(1/tan, 1/tan, 1) * Transform(persp)
    const float zn = 0.1f, zf = 10000.f;
    M4 persp = ident();
    persp.m[1][1] = -1.f;
    persp.m[2][2] = (zn + zf) / (zf - zn); persp.m[2][3] = 2 * zf * zn / (zf - zn);
    persp.m[3][2] = -1.f; persp.m[3][3] = 0.f;
    const float inv_tan = 1.0f / tanf(fov / 360.0f * pi);
    const T4 proj = mul(scale_t(inv_tan, inv_tan, 1.f), t4(persp));

    const T4 w2ns = mul(proj, w2c);
    const T4 ns2w = t4(w2ns.inv, w2ns.m);                                 // inverse(Transform)
    const T4 w2r = mul(mul(scale_t(xres * 0.5f, yres * 0.5f, 0.f), translate_t(1.0f, 1.0f, 0.0f)), w2ns);
    const T4 r2w = mul(mul(ns2w, translate_t(-1.0f, -1.0f, 0.0f)), scale_t(2.0f / xres, 2.0f / yres, 0.f));

    for (int a = 0; a < 3; a++) { out->pos[a] = pos[a]; out->forward[a] = fwd[a]; }
    out->x_res = xres; out->y_res = yres;
    const float tan_half = tanf(fov * pi / 360.0f);
    out->image_plane_dist = xres / (2.0f * tan_half);
    memcpy(out->raster_to_world, r2w.m.m, sizeof(float) * 16);
    memcpy(out->world_to_raster, w2r.m.m, sizeof(float) * 16);
}

void make_ray(const float* q, wrt_ray* r)  // Ray(origin, dir), ray.h:14-16 + Vector3::normalize vector.h:62-66
{
    float len = std::sqrt(q[3] * q[3] + q[4] * q[4] + q[5] * q[5]);
    r->ox = q[0]; r->oy = q[1]; r->oz = q[2];
    r->dx = q[3] / len; r->dy = q[4] / len; r->dz = q[5] / len;
    r->tmin = 0.f; r->tmax = kInf;
}

void camera_generate_ray(const wrt_camera& cam, float x, float y, wrt_ray* out)  // camera.cpp:37-42
{
    float raster[3] = { x, y, 0.f }, p[3];
    t_point(cam.raster_to_world, raster, p);
    float od[6] = { cam.pos[0], cam.pos[1], cam.pos[2], p[0] - cam.pos[0], p[1] - cam.pos[1], p[2] - cam.pos[2] };
    make_ray(od, out);
}

// ------------------------------------------------------------------------------------------------
// film output: ImageFilm::outputImage(filename, scale, gamma), film.cpp:39-64, color.h:47-75
// ------------------------------------------------------------------------------------------------
bool film_write(const char* path, const float* film, int w, int h, float scale, float gamma, std::string& err)
{
    std::vector<unsigned char> rgb((size_t)w * h * 3);
    const float inv_gamma = 1.f / gamma;
    for (size_t i = 0; i < (size_t)w * h * 3; i++) {
        float c = film[i] * scale;
        c = std::min(1.0f, std::max(c, 0.0f));
        c = std::pow(c, inv_gamma);
        rgb[i] = (unsigned char)(c * 255.0);
    }
    FILE* f = fopen(path, "wb");
    if (!f) { err = std::string("cannot write ") + path; return false; }
    size_t len = strlen(path);
    if (len > 4 && strcmp(path + len - 4, ".bmp") == 0) {
        const int row = (3 * w + 3) & ~3;
        unsigned char hdr[54]; memset(hdr, 0, sizeof hdr);
        unsigned size = 54 + (unsigned)row * h;
        hdr[0] = 'B'; hdr[1] = 'M'; memcpy(hdr + 2, &size, 4); hdr[10] = 54; hdr[14] = 40;
        memcpy(hdr + 18, &w, 4); memcpy(hdr + 22, &h, 4); hdr[26] = 1; hdr[28] = 24;
        fwrite(hdr, 1, 54, f);
        std::vector<unsigned char> line(row, 0);
        for (int y = h - 1; y >= 0; y--) {
            for (int x = 0; x < w; x++) {
                const unsigned char* p = &rgb[3 * ((size_t)y * w + x)];
                line[3 * x] = p[2]; line[3 * x + 1] = p[1]; line[3 * x + 2] = p[0];
            }
            fwrite(line.data(), 1, row, f);
        }
    } else {
        fprintf(f, "P6\n%d %d\n255\n", w, h);
        fwrite(rgb.data(), 1, rgb.size(), f);
    }
    fclose(f);
    return true;
}

// ------------------------------------------------------------------------------------------------
// flattened-scene cache
// ------------------------------------------------------------------------------------------------
namespace {
// Format 02: magic, then {u32 layout version, u32 sizeof(wrt_camera), u64 content fingerprint}, then the arrays.
// The fingerprint (FNV-1a over the primitive, material and light arrays) is re-checked on load: a file whose
// geometry section was damaged, or a reader built with another struct layout, is refused instead of rendered.
const char kMagic[8] = { 'W', 'R', 'T', 'S', 'C', 'N', '0', '2' };
const uint32_t kLayoutVersion = 2;

uint64_t fnv1a(uint64_t h, const void* p, size_t n)
{
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
uint64_t scene_fingerprint(const HostScene& hs)
{
    uint64_t h = 1469598103934665603ull;
    h = fnv1a(h, hs.prim_kind.data(), hs.prim_kind.size() * sizeof(int32_t));
    h = fnv1a(h, hs.prim_data.data(), hs.prim_data.size() * sizeof(float));
    h = fnv1a(h, hs.prim_matid.data(), hs.prim_matid.size() * sizeof(int32_t));
    h = fnv1a(h, hs.materials.data(), hs.materials.size() * sizeof(float));
    h = fnv1a(h, hs.lights.data(), hs.lights.size() * sizeof(float));
    return h;
}
template <class T> void put(FILE* f, const std::vector<T>& v)
{
    uint64_t n = v.size();
    fwrite(&n, sizeof n, 1, f);
    if (n) fwrite(v.data(), sizeof(T), n, f);
}
// The element count comes from the file: bound it by what is left of the file BEFORE resizing.
template <class T> bool get(FILE* f, long file_size, std::vector<T>& v)
{
    uint64_t n = 0;
    if (fread(&n, sizeof n, 1, f) != 1) return false;
    const long at = ftell(f);
    if (at < 0 || n > (uint64_t)(file_size - at) / sizeof(T)) return false;
    v.resize((size_t)n);
    return n == 0 || fread(v.data(), sizeof(T), n, f) == n;
}
}  // namespace

bool save_cache(const HostScene& hs, const char* path, std::string& err)
{
    FILE* f = fopen(path, "wb");
    if (!f) { err = std::string("cannot write ") + path; return false; }
    fwrite(kMagic, 1, 8, f);
    const uint32_t ver[2] = { kLayoutVersion, (uint32_t)sizeof(wrt_camera) };
    fwrite(ver, sizeof ver, 1, f);
    const uint64_t fp = scene_fingerprint(hs);
    fwrite(&fp, sizeof fp, 1, f);
    put(f, hs.prim_kind); put(f, hs.prim_data); put(f, hs.prim_matid); put(f, hs.materials); put(f, hs.lights);
    int32_t flags[4] = { hs.has_camera, hs.tree_built, hs.tree.dep_max, hs.tree.depth };
    fwrite(flags, sizeof flags, 1, f);
    fwrite(hs.cam_args, sizeof hs.cam_args, 1, f);
    fwrite(&hs.camera, sizeof hs.camera, 1, f);
    fwrite(hs.scene_sphere, sizeof hs.scene_sphere, 1, f);
    fwrite(hs.tree.root_box, sizeof hs.tree.root_box, 1, f);
    put(f, hs.tree.axis); put(f, hs.tree.split); put(f, hs.tree.left); put(f, hs.tree.right);
    put(f, hs.tree.first_ref); put(f, hs.tree.n_ref); put(f, hs.tree.refs);
    const bool ok = fflush(f) == 0 && !ferror(f);
    fclose(f);
    if (!ok) { err = std::string("write error on ") + path; return false; }
    return true;
}

bool load_cache(const char* path, HostScene& hs, std::string& err)
{
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    char magic[8];
    uint32_t ver[2] = { 0, 0 };
    uint64_t fp = 0;
    bool ok = size > 0 && fread(magic, 1, 8, f) == 8 && memcmp(magic, kMagic, 8) == 0;
    ok = ok && fread(ver, sizeof ver, 1, f) == 1 && ver[0] == kLayoutVersion && ver[1] == (uint32_t)sizeof(wrt_camera);
    ok = ok && fread(&fp, sizeof fp, 1, f) == 1;
    ok = ok && get(f, size, hs.prim_kind) && get(f, size, hs.prim_data) && get(f, size, hs.prim_matid) &&
         get(f, size, hs.materials) && get(f, size, hs.lights);
    int32_t flags[4] = { 0, 0, 0, 0 };
    ok = ok && fread(flags, sizeof flags, 1, f) == 1;
    ok = ok && fread(hs.cam_args, sizeof hs.cam_args, 1, f) == 1;
    ok = ok && fread(&hs.camera, sizeof hs.camera, 1, f) == 1;
    ok = ok && fread(hs.scene_sphere, sizeof hs.scene_sphere, 1, f) == 1;
    ok = ok && fread(hs.tree.root_box, sizeof hs.tree.root_box, 1, f) == 1;
    ok = ok && get(f, size, hs.tree.axis) && get(f, size, hs.tree.split) && get(f, size, hs.tree.left) && get(f, size, hs.tree.right);
    ok = ok && get(f, size, hs.tree.first_ref) && get(f, size, hs.tree.n_ref) && get(f, size, hs.tree.refs);
    fclose(f);
    if (!ok) { err = std::string("not a wrt scene cache of this build (wrong magic / layout version, or truncated): ") + path; return false; }
    // cross-array consistency: everything wrt_host_scene_desc / build_layout will index
    const size_t np = hs.prim_kind.size(), nn = hs.tree.axis.size();
    if (hs.prim_data.size() != 9 * np || hs.prim_matid.size() != np || hs.materials.size() % 11 != 0 || hs.lights.size() % 12 != 0 ||
        hs.tree.split.size() != nn || hs.tree.left.size() != nn || hs.tree.right.size() != nn ||
        hs.tree.first_ref.size() != nn || hs.tree.n_ref.size() != nn || np > (size_t)INT32_MAX || nn > (size_t)INT32_MAX) {
        err = std::string("inconsistent array lengths in scene cache: ") + path; return false;
    }
    if (scene_fingerprint(hs) != fp) { err = std::string("scene cache fingerprint mismatch (file damaged): ") + path; return false; }
    hs.has_camera = flags[0] != 0; hs.tree_built = flags[1] != 0 && nn > 0; hs.tree.dep_max = flags[2]; hs.tree.depth = flags[3];
    return true;
}

}  // namespace wrt
