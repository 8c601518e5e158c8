"""In-tree build of the native pieces (no JIT cache: the built .so files travel with the repo snapshot to the GPU box).

    lagrange_b200/lib/libwn_b200.so          nvcc, sm_100a only: all CUDA kernels + the C-ABI (include/wn_b200.h)
    lagrange_b200/lib/liblagrange_winding.so g++: the C++ host layer (lagrange::winding::FastWindingNumber over the C-ABI)
    tests/cpp/test_fast_winding_number       g++: C++ test/benchmark mirroring the reference's Catch2 file
"""
from __future__ import annotations

import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "lagrange_b200", "csrc")
LIBDIR = os.path.join(ROOT, "lagrange_b200", "lib")
LIB_CUDA = os.path.join(LIBDIR, "libwn_b200.so")
LIB_HOST = os.path.join(LIBDIR, "liblagrange_winding.so")
CPP_TEST = os.path.join(ROOT, "tests", "cpp", "test_fast_winding_number")
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-ccbin", HOST_CXX,
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-diag-suppress", "177",
]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_cuda(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "wn_b200.h")]
    if force or _newer(LIB_CUDA, deps):
        os.makedirs(LIBDIR, exist_ok=True)
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_CUDA, os.path.join(CSRC, "wn_capi.cu")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stderr)
    return LIB_CUDA


def build_host(force=False):
    src = os.path.join(ROOT, "src", "FastWindingNumber.cpp")
    hdrs = [os.path.join(ROOT, "include", "lagrange", "winding", "FastWindingNumber.h"),
            os.path.join(ROOT, "include", "lagrange", "SurfaceMesh.h"), os.path.join(ROOT, "include", "wn_b200.h"),
            os.path.join(CSRC, "wn_device.cuh"), os.path.join(CSRC, "wn_packed.h")]
    if not os.path.exists(src):
        return None
    if force or _newer(LIB_HOST, [src, LIB_CUDA] + hdrs):
        # x86-64-v3 (AVX2 + FMA: any host that carries a B200 has it) so that the single-point host traversal's fmaf() is one
        # instruction; no contraction, so its accept test rounds like the device's unfused one
        cmd = [HOST_CXX, "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "include"),
               "-o", LIB_HOST, src,
               "-L", LIBDIR, "-lwn_b200", "-Wl,-rpath,$ORIGIN"]
        subprocess.run(cmd, check=True)
    test_src = os.path.join(ROOT, "tests", "cpp", "test_fast_winding_number.cpp")
    if os.path.exists(test_src) and (force or _newer(CPP_TEST, [test_src, LIB_HOST] + hdrs)):
        cmd = [HOST_CXX, "-O2", "-std=c++17", "-Wall", "-pthread", "-I", os.path.join(ROOT, "include"), "-o", CPP_TEST, test_src, "-L", LIBDIR,
               "-llagrange_winding", "-lwn_b200", "-Wl,-rpath," + LIBDIR]
        subprocess.run(cmd, check=True)
    return LIB_HOST


def build_all(force=False, verbose=False):
    build_cuda(force, verbose)
    build_host(force)


if __name__ == "__main__":
    import sys

    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB_CUDA)
