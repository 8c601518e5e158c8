"""CPU oracle for the FastWindingNumber hot path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference's arithmetic lives in an un-vendored third-party library (UT_SolidAngle from
jdumas/WindingNumber@a48b8f5, cmake/recipes/external/winding_number.cmake:21-26) and no reference test pins a value
(modules/winding/tests/test_fast_winding_number.cpp:80-83). See oracle/wn_oracle.cpp for the restatement and its
citations.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package. The product (``lagrange_b200``) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_wn.so")
_lib = None
loaded_path = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_u64p = ctypes.POINTER(ctypes.c_uint64)

N_FLOATS_PER_LANE = 23
BVH_N = 4
CHILD_EMPTY = -1


def build(force: bool = False) -> str:
    """Compile oracle/wn_oracle.cpp (g++, OpenMP) if the shared library is missing or stale."""
    src = os.path.join(_HERE, "wn_oracle.cpp")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


def build_native() -> str | None:
    """-march=native build of the same source into oracle/_native/ (git-ignored), for the CPU arm of bench.py on the box it runs
    on (the CPU baseline is built for the host it is timed on). Same flags otherwise (-ffp-contract=off), so
    results are unchanged. Returns None if the compiler is unavailable."""
    src = os.path.join(_HERE, "wn_oracle.cpp")
    out_dir = os.path.join(_HERE, "_native")
    # one file per CPU feature set: a library built -march=native on one host must never be loaded on another (the directory
    # travels with the repo snapshot to the GPU box, whose CPU may differ)
    try:
        flags = next(ln for ln in open("/proc/cpuinfo") if ln.startswith("flags"))
    except Exception:
        flags = "unknown"
    import hashlib

    out = os.path.join(out_dir, "liboracle_wn_native_%s.so" % hashlib.sha1(flags.encode()).hexdigest()[:12])
    try:
        os.makedirs(out_dir, exist_ok=True)
        if (not os.path.exists(out)) or os.path.getmtime(out) < os.path.getmtime(src):
            subprocess.run(["/usr/bin/g++", "-O3", "-march=native", "-ffp-contract=off", "-fopenmp", "-fPIC", "-std=c++17", "-shared", "-o", out, src],
                           check=True, capture_output=True, timeout=300)
        return out
    except Exception:
        return None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        path = _LIB_PATH
        if os.environ.get("WN_ORACLE_NATIVE", "0") == "1":
            path = build_native() or _LIB_PATH
        global loaded_path
        loaded_path = path
        L = ctypes.CDLL(path)
        L.wno_num_threads.restype = ctypes.c_int
        L.wno_set_simd_lanes.argtypes = [ctypes.c_int]
        L.wno_exact64.argtypes = [_f32p, ctypes.c_int64, _i32p, ctypes.c_int64, _f32p, ctypes.c_int64, _f64p, ctypes.c_int]
        L.wno_distance64.argtypes = [_f32p, ctypes.c_int64, _i32p, ctypes.c_int64, _f32p, ctypes.c_int64, _f64p, ctypes.c_int]
        L.wno_exact32.argtypes = [_f32p, ctypes.c_int64, _i32p, ctypes.c_int64, _f32p, ctypes.c_int64, _f32p, ctypes.c_int]
        L.wno_ref_create.restype = ctypes.c_void_p
        L.wno_ref_create.argtypes = [_f32p, ctypes.c_int64, _i32p, ctypes.c_int64, ctypes.c_int]
        L.wno_ref_destroy.argtypes = [ctypes.c_void_p]
        L.wno_ref_num_nodes.restype = ctypes.c_int64
        L.wno_ref_num_nodes.argtypes = [ctypes.c_void_p]
        L.wno_ref_build_seconds.restype = ctypes.c_double
        L.wno_ref_build_seconds.argtypes = [ctypes.c_void_p]
        L.wno_ref_get_topology.argtypes = [ctypes.c_void_p, _i32p]
        L.wno_ref_get_boxdata.argtypes = [ctypes.c_void_p, _f32p]
        L.wno_ref_solid_angle.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int64, ctypes.c_float, _f32p, _u64p, ctypes.c_int]
        L.wno_ref_is_inside.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int64, ctypes.c_float, _u8p, ctypes.c_int]
        L.wno_ref_is_inside_grid.argtypes = [ctypes.c_void_p, _f32p, _f32p, _i64p, ctypes.c_int64, ctypes.c_int64,
                                             ctypes.c_int64, ctypes.c_float, _u8p, _f32p, ctypes.c_int]
        L.wno_inside_predicate.restype = ctypes.c_int
        L.wno_inside_predicate.argtypes = [ctypes.c_float]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(t)


def set_simd_lanes(on: bool) -> None:
    """Evaluate a node's 4 child lanes 4-wide with SSE (default, like upstream's v4uf path) or in a scalar loop (bit-identical)."""
    lib().wno_set_simd_lanes(1 if on else 0)


def num_threads() -> int:
    return int(lib().wno_num_threads())


def exact64(vertices, facets, queries, nthreads: int = 0) -> np.ndarray:
    """Ground truth: double-precision sum of exact triangle solid angles. Returns Omega, float64."""
    v, f, q = _f32(vertices), _i32(facets), _f32(queries).reshape(-1, 3)
    out = np.empty(len(q), dtype=np.float64)
    lib().wno_exact64(_p(v, _f32p), len(v), _p(f, _i32p), len(f), _p(q, _f32p), len(q), _p(out, _f64p), nthreads)
    return out


def distance64(vertices, facets, queries, nthreads: int = 0) -> np.ndarray:
    """Ground truth for the narrow-band distance: double-precision brute-force distance to the closest triangle."""
    v, f, q = _f32(vertices), _i32(facets), _f32(queries).reshape(-1, 3)
    out = np.empty(len(q), dtype=np.float64)
    lib().wno_distance64(_p(v, _f32p), len(v), _p(f, _i32p), len(f), _p(q, _f32p), len(q), _p(out, _f64p), nthreads)
    return out


def exact32(vertices, facets, queries, nthreads: int = 0) -> np.ndarray:
    """float32 brute force with the reference's triangle formula (the CUDA exact mode's arithmetic twin)."""
    v, f, q = _f32(vertices), _i32(facets), _f32(queries).reshape(-1, 3)
    out = np.empty(len(q), dtype=np.float32)
    lib().wno_exact32(_p(v, _f32p), len(v), _p(f, _i32p), len(f), _p(q, _f32p), len(q), _p(out, _f32p), nthreads)
    return out


def inside_predicate(omega: float) -> bool:
    """modules/winding/src/FastWindingNumber.cpp:66 evaluated exactly as written (float / double > float)."""
    return bool(lib().wno_inside_predicate(ctypes.c_float(omega)))


class RefEngine:
    """Restatement of the reference engine (UT_SolidAngle<float,float>, order 2, UT_BVH<4> BOX_AREA build)."""

    def __init__(self, vertices, facets, order: int = 2):
        self._v, self._f = _f32(vertices).reshape(-1, 3), _i32(facets).reshape(-1, 3)
        self._h = lib().wno_ref_create(_p(self._v, _f32p), len(self._v), _p(self._f, _i32p), len(self._f), order)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().wno_ref_destroy(self._h)
            self._h = None

    @property
    def num_nodes(self) -> int:
        return int(lib().wno_ref_num_nodes(self._h))

    @property
    def build_seconds(self) -> float:
        return float(lib().wno_ref_build_seconds(self._h))

    def topology(self) -> np.ndarray:
        """(n_nodes, 4) int32: >=0 internal node, -1 empty, <=-2 triangle -(c+2). Root is node 0."""
        out = np.empty((self.num_nodes, BVH_N), dtype=np.int32)
        if self.num_nodes:
            lib().wno_ref_get_topology(self._h, _p(out, _i32p))
        return out

    def boxdata(self) -> np.ndarray:
        """(n_nodes, 4, 23) float32: the stored per-child-lane values (SURVEY.md A.4 order)."""
        out = np.empty((self.num_nodes, BVH_N, N_FLOATS_PER_LANE), dtype=np.float32)
        if self.num_nodes:
            lib().wno_ref_get_boxdata(self._h, _p(out, _f32p))
        return out

    def solid_angle(self, queries, beta: float = 2.0, counters: bool = False, nthreads: int = 0):
        q = _f32(queries).reshape(-1, 3)
        out = np.empty(len(q), dtype=np.float32)
        cnt = np.zeros(3, dtype=np.uint64)
        lib().wno_ref_solid_angle(self._h, _p(q, _f32p), len(q), beta, _p(out, _f32p),
                                  _p(cnt, _u64p) if counters else None, nthreads)
        return (out, cnt) if counters else out

    def is_inside(self, queries, beta: float = 2.0, nthreads: int = 0) -> np.ndarray:
        q = _f32(queries).reshape(-1, 3)
        out = np.empty(len(q), dtype=np.uint8)
        lib().wno_ref_is_inside(self._h, _p(q, _f32p), len(q), beta, _p(out, _u8p), nthreads)
        return out

    def grid(self, origin, spacing, dims, first: int = 0, count: int | None = None, stride: int = 1, beta: float = 2.0,
             want_omega: bool = False, nthreads: int = 0):
        """Cell-centred lattice p = origin + spacing*(ijk+0.5), x fastest (mesh_to_volume.cpp:147-149)."""
        o, s = _f32(origin), _f32(spacing)
        d = np.ascontiguousarray(dims, dtype=np.int64)
        total = int(d[0] * d[1] * d[2])
        if count is None:
            count = (total - first + stride - 1) // stride
        inside = np.empty(count, dtype=np.uint8)
        omega = np.empty(count, dtype=np.float32) if want_omega else None
        lib().wno_ref_is_inside_grid(self._h, _p(o, _f32p), _p(s, _f32p), _p(d, _i64p), first, count, stride, beta,
                                     _p(inside, _u8p), _p(omega, _f32p) if want_omega else None, nthreads)
        return (inside, omega) if want_omega else inside
