"""In-tree build of libwrt_b200.so (sm_100a) — the product library behind include/wrt.h.

    python winmad-s-raytracer-v1.0_b200/build.py [--force] [--hostsim]

CUDA sources are compiled with nvcc for sm_100a only (-gencode arch=compute_100a,code=sm_100a),
-lineinfo for ncu source pages, and -fmad=false: the traversal/intersection arithmetic must not be
contracted into FMAs or hit distances stop being bit-identical to the reference's x86 results.
Host sources are compiled with g++ -ffp-contract=off for the same reason (the KD build must
reproduce the reference tree exactly).  No GPU is needed to build.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libwrt_b200.so")

CU_SOURCES = ["csrc/scene_upload.cu", "csrc/trace_kernels.cu", "csrc/pt_wavefront.cu",
              "csrc/bdpt_wavefront.cu", "csrc/debug_kernels.cu", "csrc/multi_gpu.cu"]
CPP_SOURCES = ["csrc/scene_layout.cpp", "host/kd_build.cpp", "host/scene_io.cpp", "host/host_api.cpp"]
HEADERS = ["csrc/dev_scene.h", "csrc/hd_compat.h", "csrc/traverse.cuh", "csrc/scene_layout.h", "csrc/pt_logic.cuh", "csrc/whitted_logic.cuh",
           "csrc/bdpt_logic.cuh", "csrc/warp_utils.cuh", "csrc/wavefront_kernels.cuh", "csrc/trace_persistent.cuh", "csrc/trace_pooled.cuh",
           "csrc/shading.cuh", "csrc/shading_kat.cuh", "csrc/wavefront.h", "host/host_scene.h", "../include/wrt.h"]

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = os.environ.get("CXX") or shutil.which("g++") or "g++"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false",
              "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off", "-diag-suppress", "177"]
CXX_FLAGS = ["-O2", "-ffp-contract=off", "-fPIC", "-std=c++17", "-Wall"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    print("[wrt build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, cwd=HERE)


def build(force=False, verbose_ptxas=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(HERE, h) for h in HEADERS]
    objs = []
    for src in CU_SOURCES:
        o = os.path.join(OBJ, os.path.basename(src) + ".o")
        if force or _newer(o, [os.path.join(HERE, src), os.path.abspath(__file__)] + hdrs):
            _run([NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose_ptxas else []) + ["-c", src, "-o", o])
        objs.append(o)
    for src in CPP_SOURCES:
        o = os.path.join(OBJ, os.path.basename(src) + ".o")
        if force or _newer(o, [os.path.join(HERE, src), os.path.abspath(__file__)] + hdrs):
            _run([CXX] + CXX_FLAGS + ["-I", os.path.dirname(NVCC) + "/../include", "-c", src, "-o", o])
        objs.append(o)
    if force or _newer(LIB, objs):
        _run([NVCC, "-shared", "-o", LIB] + objs)  # nvcc links the static cudart by default
    cli = os.path.join(HERE, "wrt_tot")
    if force or _newer(cli, [os.path.join(HERE, "host/main.cpp"), LIB]):
        _run([CXX] + CXX_FLAGS + ["-o", cli, "host/main.cpp", "-L", HERE, "-lwrt_b200", "-Wl,-rpath," + HERE])
    return LIB


def build_hostsim(force=False):
    """Test-only library: the same __host__ __device__ per-ray / per-path code compiled for the CPU.
    Lives under tests/hostsim/, is never loaded by the product package."""
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    out = os.path.join(ROOT, "tests", "hostsim", "libwrt_hostsim.so")
    deps = [src, os.path.join(HERE, "csrc/scene_layout.cpp")] + [os.path.join(HERE, h) for h in HEADERS]
    if force or _newer(out, deps):
        _run([CXX, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-Wall",
              "-Wno-unknown-pragmas", "-DWRT_HOSTSIM", "-I", os.path.join(HERE, "csrc"), "-o", out, src,
              os.path.join(HERE, "csrc/scene_layout.cpp")])
    return out


def build_warpsim(force=False):
    """Test-only library: the warp-level schedulers (csrc/trace_pooled.cuh, trace_persistent.cuh) compiled for the CPU on top
    of an emulation of CUDA's warp primitives (tests/hostsim/warpsim.cpp).  Never loaded by the product package."""
    src = os.path.join(ROOT, "tests", "hostsim", "warpsim.cpp")
    out = os.path.join(ROOT, "tests", "hostsim", "libwrt_warpsim.so")
    deps = [src, os.path.join(HERE, "csrc/scene_layout.cpp")] + [os.path.join(HERE, h) for h in HEADERS]
    extra = os.environ.get("WRT_WARPSIM_DEFS", "").split()      # e.g. -DWRT_STACK8=0 to test a non-default scheduler variant
    if force or extra or _newer(out, deps):
        _run([CXX, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-Wall", "-Wno-unknown-pragmas",
              "-Wno-unused-function", "-Wno-unused-variable", "-DWRT_HOSTSIM", "-DWRT_WARPSIM"] + extra + ["-I", os.path.join(HERE, "csrc"),
              "-o", out, src, os.path.join(HERE, "csrc/scene_layout.cpp"), "-lpthread"])
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose_ptxas="--ptxas" in sys.argv)
    if "--hostsim" in sys.argv:
        build_hostsim(force="--force" in sys.argv)
