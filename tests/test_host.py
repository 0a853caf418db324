"""Host logic (no GPU): the KD-tree builder reproduces the reference tree exactly, the .scene/OBJ
loader reproduces Scene::loadScene, camera set-up, the C-ABI library loads and exports every symbol
include/wrt.h declares, argument checking / error behaviour."""
import os
import re

import numpy as np
import pytest

import scenes
import util

FIXTURES = ["torus", "cbox_dragon", "bunny", "small_mixed", "mixed_torus"]


@pytest.mark.parametrize("name", FIXTURES)
def test_kd_build_matches_reference_digest(wrt, name):
    """T2: our builder's flattened tree == the reference's (digest stored by make_golden.py)."""
    sc, z = scenes.load_fixture(name)
    hs = util.host_scene(wrt, sc)
    t = hs.arrays()["tree"]
    assert len(t["axis"]) == int(z["tree_nodes"]) and len(t["refs"]) == int(z["tree_refs"])
    assert util.tree_digest(t) == str(z["tree_sha"])
    assert np.array_equal(util.bits(t["root_box"]), util.bits(z["root_box"]))
    assert np.array_equal(util.bits(hs.scene_sphere()), util.bits(z["scene_sphere"]))


@pytest.mark.parametrize("make", [lambda: scenes.small_mixed_scene(), lambda: scenes.cornell_box_scene(),
                                  lambda: scenes.synthetic_torus_scene(n=48, width=64, height=64, n_spheres=300)])
def test_kd_build_matches_reference_live(wrt, have_ref, make):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    sc = make()
    hs = util.host_scene(wrt, sc)
    ref = util.ref_scene(sc)
    a, b = hs.arrays()["tree"], ref.tree()
    for k in ("axis", "left", "right", "nref", "refs"):
        assert np.array_equal(a[k], b[k]), k
    inner = b["axis"] >= 0
    assert np.array_equal(util.bits(a["split"][inner]), util.bits(b["split"][inner]))
    assert np.array_equal(a["first_ref"][~inner], b["first_ref"][~inner])


def test_parallel_kd_build_equals_serial(wrt, monkeypatch):
    """The multi-threaded build (concurrent axis sorts, per-axis event distribution, sub-tree tasks spliced back in
    DFS pre-order) must emit the same arrays, index for index, as the serial build (WRT_KD_THREADS=1) — large enough
    to take every parallel path (> 200 000 events per axis, > 4096 primitives per task)."""
    sc = scenes.synthetic_torus_scene(n=240, width=64, height=64, n_spheres=2000)
    monkeypatch.setenv("WRT_KD_THREADS", "1")
    a = util.host_scene(wrt, sc).arrays()["tree"]
    monkeypatch.setenv("WRT_KD_CHUNK_MIN", "50000")      # also split the per-axis event distribution of the top nodes into ranges
    for threads in ("2", "7", "64"):
        monkeypatch.setenv("WRT_KD_THREADS", threads)
        b = util.host_scene(wrt, sc).arrays()["tree"]
        for k in ("axis", "left", "right", "first_ref", "nref", "refs"):
            assert np.array_equal(a[k], b[k]), (threads, k)
        assert np.array_equal(util.bits(a["split"]), util.bits(b["split"]))
        assert np.array_equal(util.bits(a["root_box"]), util.bits(b["root_box"]))


def test_parallel_layout_equals_serial(wrt, monkeypatch):
    """build_layout (csrc/scene_layout.cpp) on threads writes the same bytes as on one: node array, leaf records incl. skip records,
    primitive table — on a mesh + spheres scene with every leaf >= 3 entries chunked."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    from hostsim_py import HostSim
    monkeypatch.setenv("WRT_LEAF_SKIP_MIN", "3"); monkeypatch.setenv("WRT_LEAF_SKIP_CHUNK", "2")
    sc = scenes.synthetic_torus_scene(n=160, width=64, height=64, n_spheres=3000)
    hs = util.host_scene(wrt, sc)
    monkeypatch.setenv("WRT_LAYOUT_THREADS", "1")
    a = HostSim(hs.desc(), hs)
    want, nrec = a.layout_digest(), a.num_recs()
    for t in ("2", "7", "16"):
        monkeypatch.setenv("WRT_LAYOUT_THREADS", t)
        b = HostSim(hs.desc(), hs)
        assert b.num_recs() == nrec and b.layout_digest() == want, t


def _eps_dense_scene(n=64, scale=0.004):
    """A mesh whose box bounds lie closer together than the reference's EPS (1e-3): the displaced torus scaled down until a triangle is
    ~1e-4 across.  Long chains of event positions are then pairwise "equal" under cmp(), the comparator is as non-transitive as it gets,
    and the sorted order is whatever the sorting ALGORITHM makes of it — the case the exact restatement of glibc's merge sort exists for."""
    sc = scenes.synthetic_torus_scene(n=n, width=64, height=64, n_spheres=0)
    sc.data = (sc.data.astype(np.float64) * scale).astype(np.float32)
    return sc


def test_parallel_sort_and_split_search_on_epsilon_dense_events(wrt, have_ref, monkeypatch):
    """Events denser than EPS: the multi-threaded build (merge-sort restatement on threads, findSplitPlane over per-range prefix-minima
    staircases, counted distribution) == the serial build with libc qsort == the reference's own tree."""
    sc = _eps_dense_scene()
    ext = sc.data.reshape(-1, 3)[: -6]
    assert np.ptp(ext[:, 0]) / len(sc.data) ** 0.5 < 1e-3            # mean spacing of the bounds well below EPS
    monkeypatch.setenv("WRT_KD_THREADS", "1")
    a = util.host_scene(wrt, sc).arrays()["tree"]
    monkeypatch.setenv("WRT_KD_CHUNK_MIN", "4000"); monkeypatch.setenv("WRT_KD_SORT_MIN", "1000")
    for threads, sort in (("8", ""), ("5", ""), ("8", "libc")):
        monkeypatch.setenv("WRT_KD_THREADS", threads); monkeypatch.setenv("WRT_KD_SORT", sort)
        b = util.host_scene(wrt, sc).arrays()["tree"]
        for k in ("axis", "left", "right", "first_ref", "nref", "refs"):
            assert np.array_equal(a[k], b[k]), (threads, sort, k)
        assert np.array_equal(util.bits(a["split"]), util.bits(b["split"]))
    if have_ref:
        r = util.ref_scene(sc).tree()
        for k in ("axis", "left", "right", "nref", "refs"):
            assert np.array_equal(a[k], r[k]), k
        inner = r["axis"] >= 0
        assert np.array_equal(util.bits(a["split"][inner]), util.bits(r["split"][inner]))


def test_scene_loader_matches_reference(wrt, have_ref):
    """Our XML/OBJ loader vs Scene::loadScene on torus.scene (needs /root/reference: build container only)."""
    from oracle import refpy
    if not have_ref or not os.path.isdir(refpy.REF_ROOT):
        pytest.skip("reference tree not present")
    os.environ["WRT_OBJ_DIR"] = refpy.REF_ROOT + "/ObjFiles"
    hs = wrt.HostScene.load(os.path.join(refpy.REF_ROOT, "torus.scene"))   # Windows paths resolved by base name
    sc, z = scenes.load_fixture("torus")
    a = hs.arrays()
    assert np.array_equal(a["prim_kind"], sc.kind) and np.array_equal(a["prim_matid"], sc.matid)
    assert np.array_equal(util.bits(a["prim_data"]), util.bits(sc.data))
    assert np.array_equal(util.bits(a["materials"]), util.bits(sc.materials))
    assert np.array_equal(util.bits(a["lights"]), util.bits(sc.lights))
    assert util.tree_digest(a["tree"]) == str(z["tree_sha"])
    cam = hs.camera()
    c45 = z["cam45"]
    assert np.allclose(np.array(cam.raster_to_world[:]), c45[13:29], rtol=2e-6, atol=1e-9)
    assert np.allclose(np.array(cam.world_to_raster[:]), c45[29:45], rtol=2e-6, atol=1e-9)
    assert abs(cam.image_plane_dist - c45[12]) <= 1e-6 * abs(c45[12])


def test_scene_files_round_trip_and_match_reference_loader(wrt, have_ref, tmp_path):
    """A scene written in the reference's .scene + OBJ format loads into the same primitives, materials, lights and
    camera through our loader (host/scene_io.cpp) — and, where the compiled reference is present, through Scene::loadScene
    (same object order, bit-identical vertex data, same KD-tree)."""
    sc = scenes.cornell_box_scene(64, 64)
    path = scenes.write_scene_files(sc, str(tmp_path))
    hs = wrt.HostScene.load(path)
    a = hs.arrays()
    assert set(map(bytes, a["prim_data"].reshape(-1, 9))) == set(map(bytes, sc.data))
    assert np.array_equal(util.bits(a["lights"]), util.bits(np.asarray(sc.lights, np.float32)))
    assert np.array_equal(util.bits(a["materials"]), util.bits(np.asarray(sc.materials, np.float32)))
    if have_ref:
        from oracle import refpy
        ref = refpy.RefScene("pt")
        n = ref.load_file(path, 64, 64)
        kind, data, mat = ref.prims()
        assert n == len(a["prim_kind"]) and np.array_equal(kind, a["prim_kind"]) and np.array_equal(mat, a["prim_matid"])
        assert np.array_equal(util.bits(data), util.bits(a["prim_data"]))
        b = ref.tree()
        for k in ("axis", "left", "right", "nref", "refs"):
            assert np.array_equal(a["tree"][k], b[k]), k


@pytest.mark.parametrize("name", FIXTURES)
def test_camera_setup_close_to_reference(wrt, name):
    sc, z = scenes.load_fixture(name)
    c = sc.cam12
    cam = wrt.camera_setup(c[0:3], c[3:6], c[6:9], c[9], c[10], c[11])
    c45 = z["cam45"]
    # tolerance: matrices agree to a few ulp (different but equivalent 4x4 inversion), stated 2e-6 relative
    assert np.allclose(np.array(cam.raster_to_world[:]), c45[13:29], rtol=2e-6, atol=1e-7)
    assert np.allclose(np.array(cam.world_to_raster[:]), c45[29:45], rtol=2e-6, atol=1e-7)


def test_cabi_exports_every_declared_symbol(wrt):
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "wrt.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(wrt_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    L = wrt.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(declared) == sorted(wrt.EXPORTS)


def test_error_behaviour_without_gpu_or_bad_input(wrt):
    with pytest.raises(wrt.WrtError):
        wrt.HostScene.load("/nonexistent/file.scene")
    with pytest.raises(wrt.WrtError):   # unknown primitive kind
        wrt.HostScene.from_arrays(np.zeros((1, 11)), [7], np.zeros((1, 9)), [1], np.zeros((0, 12)), None, build=False)
    hs = wrt.HostScene.from_arrays(np.zeros((1, 11)), np.zeros(0, np.int32), np.zeros((0, 9)), np.zeros(0, np.int32),
                                   np.zeros((0, 12)), None, build=False)
    with pytest.raises(wrt.WrtError):   # empty scene: nothing to build
        hs.build_kdtree()
    if wrt.device_count() == 0:
        sc = scenes.small_mixed_scene()
        with pytest.raises(wrt.WrtError) as e:   # no CPU fallback: creating a device scene must fail loudly
            wrt.Scene(util.host_scene(wrt, sc))
        assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_cache_round_trip(wrt, tmp_path):
    sc = scenes.small_mixed_scene()
    hs = util.host_scene(wrt, sc)
    p = str(tmp_path / "s.wrtscene")
    hs.save(p)
    hs2 = wrt.HostScene.load_cache(p)
    a, b = hs.arrays(), hs2.arrays()
    for k in ("prim_kind", "prim_data", "prim_matid", "materials", "lights"):
        assert np.array_equal(a[k], b[k])
    assert util.tree_digest(a["tree"]) == util.tree_digest(b["tree"])


def test_parameters_file(wrt, tmp_path):
    p = tmp_path / "parameters.para"
    p.write_text("#MAX_TRACING_DEPTH\n7\n\n#SAMPLES_PER_PIXEL\n4\n#l\n8\n#h\n4\n#WIDTH\n640\n#HEIGHT\n480\n#x\n5\n#y\n400\n")
    para = wrt.Parameters().load_parameters(str(p))
    assert (para.MAX_TRACING_DEPTH, para.SAMPLES_PER_PIXEL, para.WIDTH, para.HEIGHT) == (7, 4, 640, 480)


def test_film_write(wrt, tmp_path):
    film = np.zeros((4, 6, 3), np.float32); film[1, 2] = [0.5, 2.0, -1.0]
    p = str(tmp_path / "o.ppm")
    wrt.film_write(p, film)
    raw = open(p, "rb").read()
    assert raw.startswith(b"P6\n6 4\n255\n")
    px = np.frombuffer(raw[len(b"P6\n6 4\n255\n"):], np.uint8).reshape(4, 6, 3)
    assert tuple(px[1, 2]) == (int(0.5 ** (1 / 2.2) * 255.0), 255, 0)


def test_integration_shim_is_the_documented_code_and_links_with_the_reference(wrt, tmp_path):
    """INTEGRATION.md's reference-side binding is real code: its C++ block is byte for byte oracle/shim_integration/gpuIntegrator.h,
    which oracle/Makefile compiles against the UNMODIFIED reference headers and links with the reference's own objects and
    libwrt_b200.so into oracle/_ref/ToT_gpu (the reference's command line + the -gp / -gr / -gbpt branches).  Here: the text
    matches, and where the binary is built its pure-reference branch (-p) renders a scene file on the CPU."""
    import re, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.findall(r"```cpp\n(.*?)```", doc, re.S)[0]
    assert block == open(os.path.join(root, "oracle", "shim_integration", "gpuIntegrator.h")).read()
    exe = os.path.join(root, "oracle", "_ref", "ToT_gpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ToT_gpu not built (needs /root/reference at build time)")
    sc = scenes.cornell_box_scene(32, 32)
    scene_file = scenes.write_scene_files(sc, str(tmp_path))
    para = tmp_path / "parameters.para"
    para.write_text("#MAX_TRACING_DEPTH\n5\n#SAMPLES_PER_PIXEL\n4\n#l\n8\n#h\n4\n#WIDTH\n32\n#HEIGHT\n32\n#x\n5\n#y\n400\n")
    r = subprocess.run([exe, scene_file, str(tmp_path / "ref.ppm"), "-p", str(para)], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and (tmp_path / "ref.ppm").stat().st_size > 32 * 32 * 3


def test_cache_rejects_damaged_files(wrt, tmp_path):
    """a truncated / corrupt / stale-format cache must fail with an error code, not terminate the process
    (absurd element counts are bounded by the file size before any allocation; arrays are cross-checked; the geometry
    fingerprint is re-computed)."""
    sc = scenes.small_mixed_scene()
    hs = util.host_scene(wrt, sc)
    p = str(tmp_path / "s.wrtscene")
    hs.save(p)
    raw = open(p, "rb").read()
    cases = {
        "truncated": raw[: len(raw) // 2],
        "huge_count": raw[:24] + (2 ** 62).to_bytes(8, "little") + raw[32:],     # first array claims 2^62 elements
        "old_magic": b"WRTSCN01" + raw[8:],
        "flipped_vertex": raw[:200] + bytes([raw[200] ^ 0x40]) + raw[201:],       # inside prim_data: fingerprint mismatch
        "empty": b"",
    }
    for name, blob in cases.items():
        q = str(tmp_path / (name + ".wrtscene"))
        open(q, "wb").write(blob)
        with pytest.raises(wrt.WrtError):
            wrt.HostScene.load_cache(q)
    assert wrt.HostScene.load_cache(p).arrays()["prim_kind"].shape == hs.arrays()["prim_kind"].shape


def test_layout_rejects_what_the_kernels_cannot_index(wrt):
    """material / light ids past their tables and trees deeper than the traversal stack are refused when the
    layout is built (the kernels index materials[matid], lights[-matid-1] and a 32-entry stack without checks)."""
    from hostsim_py import HostSim
    sc = scenes.small_mixed_scene()
    hs = util.host_scene(wrt, sc)
    a = hs.arrays()
    t = a["tree"]

    def build(matid=None, tree=None):
        d, keep = wrt.desc_from_arrays(a["prim_kind"], a["prim_data"], a["prim_matid"] if matid is None else matid,
                                       a["materials"], a["lights"], tree or t)
        return HostSim(d, keep)

    build()                                                   # the untouched scene is fine
    bad = a["prim_matid"].copy(); bad[0] = len(a["materials"])
    with pytest.raises(RuntimeError, match="material id"):
        build(matid=bad)
    bad = a["prim_matid"].copy(); bad[0] = -(len(a["lights"]) + 1)
    with pytest.raises(RuntimeError, match="light"):
        build(matid=bad)
    # a degenerate chain of 40 interior nodes above the real root
    n0 = len(t["axis"]); extra = 40
    chain = dict(t)
    axis = np.concatenate([np.zeros(extra, np.int32), t["axis"], np.full(extra, -1, np.int32)])
    split = np.concatenate([np.full(extra, -1e6, np.float32), t["split"], np.zeros(extra, np.float32)])
    # chain node i: left = empty leaf (extra + n0 + i), right = next chain node (or the old root)
    left = np.concatenate([extra + n0 + np.arange(extra, dtype=np.int32), np.where(t["left"] >= 0, t["left"] + extra, -1), np.full(extra, -1, np.int32)])
    right = np.concatenate([np.arange(1, extra + 1, dtype=np.int32), np.where(t["right"] >= 0, t["right"] + extra, -1), np.full(extra, -1, np.int32)])
    first = np.concatenate([np.full(extra, -1, np.int32), t["first_ref"], np.zeros(extra, np.int32)])
    nref = np.concatenate([np.zeros(extra, np.int32), t["nref"], np.zeros(extra, np.int32)])
    chain.update(axis=axis, split=split, left=left.astype(np.int32), right=right.astype(np.int32), first_ref=first, nref=nref)
    with pytest.raises(RuntimeError, match="too deep"):
        build(tree=chain)


def test_kd_build_matches_reference_at_headline_scale(wrt, have_ref):
    """VERDICT r1 'weak' 4: the tree the headline benchmark (C3, 1 002 528 triangles + 2 light triangles) traverses is the
    reference's own tree, node for node and reference for reference — KDtreeAccel::init + buildTree (KDtreeAccel.cpp:12-307)
    run by the compiled reference against host/kd_build.cpp, live (about 10 s)."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    sc = scenes.synthetic_torus_scene(n=708, width=1920, height=1080)
    assert sc.n_prims == 1002530
    hs = util.host_scene(wrt, sc)
    ref = util.ref_scene(sc)
    a, b = hs.arrays()["tree"], ref.tree()
    assert len(a["axis"]) == len(b["axis"]) > 300000 and len(a["refs"]) == len(b["refs"]) > 1500000
    for k in ("axis", "left", "right", "nref", "refs"):
        assert np.array_equal(a[k], b[k]), k
    inner = b["axis"] >= 0
    assert np.array_equal(util.bits(a["split"][inner]), util.bits(b["split"][inner]))
    assert np.array_equal(a["first_ref"][~inner], b["first_ref"][~inner])
    assert np.array_equal(util.bits(a["root_box"]), util.bits(b["box"][0]))


def test_camera_setup_bit_identical_to_reference(wrt, have_ref):
    """Camera::setup (camera.cpp:3-29, lookAt / perspective / inverse of transform.cpp) restated operation for operation:
    rasterToWorld, worldToRaster and imagePlaneDist are bit-identical — incl. torus.scene's camera, whose coordinates are in
    the thousands (an ulp in the matrix moves primary rays there)."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    cams = [scenes.load_fixture(n)[0].cam12 for n in FIXTURES] + [scenes.synthetic_torus_scene(n=8, width=1920, height=1080).cam12,
            np.array([3, -2, 1.5, -0.3, 0.9, -0.1, 0.1, 0.2, 0.97, 640, 480, 63.5], np.float32)]
    sc = scenes.small_mixed_scene()
    for c in cams:
        sc.cam12 = np.asarray(c, np.float32)
        want = util.ref_scene(sc).camera()
        cam = wrt.camera_setup(c[0:3], c[3:6], c[6:9], c[9], c[10], c[11])
        assert np.array_equal(util.bits(np.array(cam.raster_to_world[:], np.float32)), util.bits(want[13:29]))
        assert np.array_equal(util.bits(np.array(cam.world_to_raster[:], np.float32)), util.bits(want[29:45]))
        assert np.float32(cam.image_plane_dist) == want[12]
        assert np.array_equal(util.bits(np.array(list(cam.pos) + list(cam.forward), np.float32)), util.bits(want[0:6]))
