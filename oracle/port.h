/* TEST INFRASTRUCTURE ONLY — plain-C restatement ("port") of the reference's hot path.
 * Never linked into, imported by or executed from the product library. */
#ifndef WRT_ORACLE_PORT_H
#define WRT_ORACLE_PORT_H
#include <stdint.h>

typedef struct {
    /* Scene::objs (R/src/scene/scene.h:21) */
    int32_t n_prims;
    const int32_t* kind;      /* 0 triangle, 1 sphere */
    const float* data9;       /* triangle p0 p1 p2 | sphere c.xyz r */
    const int32_t* matid;
    /* KDtreeAccelNode tree, any order, root = 0 (R/src/scene/KDtreeAccel.h:26-45) */
    int32_t n_nodes;
    const int32_t* axis; const float* split; const int32_t* left; const int32_t* right;
    const int32_t* first_ref; const int32_t* n_ref; const int32_t* refs;
    float root_box[6];
} port_scene;

/* counters[4]: interior visits, leaf visits, triangle tests, sphere tests (may be NULL) */
void port_intersect(const port_scene* s, const float* rays8, long long n, int32_t* prim, float* t,
                    float* p3, float* n3, int32_t* inside, int32_t* matid, unsigned long long* counters);
void port_occluded(const port_scene* s, const float* p1_dir_p2, long long n, uint8_t* occluded);
void port_make_rays(const float* origin_dir6, long long n, float* rays8);
void port_shadow_test(const port_scene* s, const float* rays8, const float* target3, long long n, float* visible);
void port_intersect_any(const port_scene* s, const float* rays8, long long n, uint8_t* hit);
int port_triangle_hit(const float* tri9, const float* ray8, float* t);
int port_sphere_hit(const float* cr4, const float* ray8, float* t, int* inside);
int port_aabb_hit(const float* box6, const float* ray8, float* t1, float* t2);
#endif
