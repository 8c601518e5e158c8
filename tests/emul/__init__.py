"""Host emulation of the device build (TEST INFRASTRUCTURE ONLY): ctypes loader for tests/emul/wn_emul.cpp."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libwn_emul.so")
_SRC = os.path.join(_HERE, "wn_emul.cpp")
_CSRC = os.path.join(_HERE, "..", "..", "lagrange_b200", "csrc")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(force=False):
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in ("wn_device.cuh", "wn_build_core.cuh", "wn_refbuild_core.cuh")]
    stale = (not os.path.exists(_LIB)) or any(os.path.getmtime(_LIB) < os.path.getmtime(d) for d in deps)
    if force or stale:
        extra = os.environ.get("WN_EMUL_DEFINES", "").split()
        subprocess.run(["/usr/bin/g++", *extra, "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-fPIC", "-std=c++17", "-w", "-shared",
                        "-o", _LIB, _SRC], check=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        L.emul_build.restype = ctypes.c_void_p
        L.emul_build.argtypes = [_f32p, ctypes.c_int64, _i32p, ctypes.c_int64, _i32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.emul_destroy.argtypes = [ctypes.c_void_p]
        L.emul_ref_topology.restype = ctypes.c_int64
        L.emul_ref_topology.argtypes = [_f32p, ctypes.c_int64, _i32p, ctypes.c_int64, _i32p, _i32p, _i32p]
        for name in ("emul_error", "emul_max_depth", "emul_width"):
            getattr(L, name).restype = ctypes.c_int
            getattr(L, name).argtypes = [ctypes.c_void_p]
        for name in ("emul_num_internal", "emul_num_entries"):
            getattr(L, name).restype = ctypes.c_int64
            getattr(L, name).argtypes = [ctypes.c_void_p]
        L.emul_get_topology.argtypes = [ctypes.c_void_p, _i32p]
        L.emul_get_ref23.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, _f32p]
        L.emul_get_packed.argtypes = [ctypes.c_void_p, _f32p, _i32p, _f32p, _u32p]
        L.emul_query.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int64, ctypes.c_float, _f32p, _u64p]
        L.emul_inside_from_omega.restype = ctypes.c_int
        L.emul_inside_from_omega.argtypes = [ctypes.c_float]
        L.emul_tile_cost.argtypes = [ctypes.c_void_p, _f32p, _f32p, ctypes.POINTER(ctypes.c_int64), ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                     _i32p, _i32p, ctypes.POINTER(ctypes.c_double)]
        L.emul_point_tri_dist2.argtypes = [_f32p, _f32p, ctypes.c_int64, _f32p]
        L.emul_lattice_coord.restype = ctypes.c_float
        L.emul_lattice_coord.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_int]
        _lib = L
    return _lib


class EmulEngine:
    def __init__(self, vertices, facets, child=None, leaf_size=1, order=2, radius_mode=0, approx_single=None, morton_bits=63,
                 hierarchy="lbvh"):
        self.v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        self.f = np.ascontiguousarray(facets, dtype=np.int32).reshape(-1, 3)
        if approx_single is None:
            approx_single = 1 if child is not None else 0
        if child is not None:
            child = np.ascontiguousarray(child, dtype=np.int32)
            nn, w = child.shape
            cp = child.ctypes.data_as(_i32p)
        else:
            nn, w, cp = 0, 0, None
        self._h = lib().emul_build(self.v.ctypes.data_as(_f32p), len(self.v), self.f.ctypes.data_as(_i32p), len(self.f), cp, nn, w,
                                   leaf_size, order, radius_mode, approx_single, morton_bits, {"lbvh": 0, "kd": 1, "kd_sah": 2}[hierarchy])

    def __del__(self):
        if getattr(self, "_h", None):
            lib().emul_destroy(self._h)
            self._h = None

    error = property(lambda self: lib().emul_error(self._h))
    num_internal = property(lambda self: int(lib().emul_num_internal(self._h)))
    num_entries = property(lambda self: int(lib().emul_num_entries(self._h)))
    max_depth = property(lambda self: lib().emul_max_depth(self._h))
    width = property(lambda self: lib().emul_width(self._h))

    def topology(self):
        out = np.empty((self.num_internal, self.width), dtype=np.int32)
        lib().emul_get_topology(self._h, out.ctypes.data_as(_i32p))
        return out

    def ref23(self, first=0, count=None):
        if count is None:
            count = self.num_internal + len(self.f) - first
        out = np.empty((count, 23), dtype=np.float32)
        lib().emul_get_ref23(self._h, first, count, out.ctypes.data_as(_f32p))
        return out

    def packed(self):
        n, nt = self.num_entries, len(self.f)
        rec = np.empty((6, n, 4), dtype=np.float32)
        link = np.empty(n, dtype=np.int32)
        tris = np.empty((nt, 3, 4), dtype=np.float32)
        order = np.empty(nt, dtype=np.uint32)
        lib().emul_get_packed(self._h, rec.ctypes.data_as(_f32p), link.ctypes.data_as(_i32p), tris.ctypes.data_as(_f32p),
                              order.ctypes.data_as(_u32p))
        return rec, link, tris, order

    def tile_cost(self, origin, spacing, dims, tile_stride=4, beta=2.0, kappa=6.0, warp_shape=(4, 4, 4), tile_shape=(8, 8, 8)):
        """Cost model of the tiled query path on this tree (see emul_tile_cost). Returns a dict of per-tile averages."""
        o = np.ascontiguousarray(origin, dtype=np.float32)
        s = np.ascontiguousarray(spacing, dtype=np.float32)
        d = np.ascontiguousarray(dims, dtype=np.int64)
        out = np.zeros(8, dtype=np.float64)
        ws = np.ascontiguousarray(warp_shape, dtype=np.int32)
        ts = np.ascontiguousarray(tile_shape, dtype=np.int32)
        assert int(np.prod(ws)) in (32, 64, 128, 256) and all(int(t) % int(w) == 0 for t, w in zip(ts, ws))
        groups = int(np.prod(ws)) // 32
        lib().emul_tile_cost(self._h, o.ctypes.data_as(_f32p), s.ctypes.data_as(_f32p), d.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                             int(tile_stride), float(beta), float(kappa), ws.ctypes.data_as(_i32p), ts.ctypes.data_as(_i32p),
                             out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        tiles = max(out[0], 1.0)
        names = ("tiles", "walk_steps", "evaluations", "far_set", "direct", "conditional_items", "exact_triangle_evals", "warps")
        r = {k: (v if k in ("tiles", "warps") else v / tiles) for k, v in zip(names, out)}
        # warp-instructions per tile: the measured per-step costs of k_tile_query (45 test/control + 37 per evaluation + fixed part
        # per warp) and k_tile_plan (BFS/sort/packets + far-set sampling at 2 x 37 per record and warp + list handling)
        warps = int(np.prod(ts)) // int(np.prod(ws))
        r["query_instr"] = ((25.0 + 10.0 * groups) * r["walk_steps"] + 37.0 * r["evaluations"] + 50.0 * r["exact_triangle_evals"]
                            + warps * (37.0 * groups * r["direct"] + 150.0 + 100.0 * groups))
        r["plan_instr"] = 3000.0 + 95.0 * r["far_set"] + 12.0 * r["conditional_items"]
        r["instr_per_tile"] = r["query_instr"] + r["plan_instr"]
        r["instr_per_point"] = r["instr_per_tile"] / float(np.prod(ts))
        return r

    def solid_angle(self, queries, beta=2.0, counters=False):
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, 3)
        out = np.empty(len(q), dtype=np.float32)
        cnt = np.zeros(3, dtype=np.uint64)
        lib().emul_query(self._h, q.ctypes.data_as(_f32p), len(q), beta, out.ctypes.data_as(_f32p),
                         cnt.ctypes.data_as(_u64p) if counters else None)
        return (out, cnt) if counters else out


def point_tri_dist2(points, tris):
    """Squared distances point[i] -> triangle tris[i] ([n,3,3]) with the device's float routine (wn_point_tri_dist2)."""
    p = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9)
    out = np.empty(len(p), dtype=np.float32)
    lib().emul_point_tri_dist2(p.ctypes.data_as(_f32p), t.ctypes.data_as(_f32p), len(p), out.ctypes.data_as(_f32p))
    return out


def ref_topology(vertices, facets):
    """K3R (WN_HIERARCHY_REFERENCE) on the host: the driver of lagrange_b200/csrc/wn_refbuild_core.cuh with a sequential
    backend. Returns (child table [n_nodes, 4] int32, levels, host synchronisations the GPU backend would make)."""
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    f = np.ascontiguousarray(facets, dtype=np.int32).reshape(-1, 3)
    out = np.full((max(1, len(f)), 4), -1, dtype=np.int32)
    levels, syncs = ctypes.c_int32(0), ctypes.c_int32(0)
    n = lib().emul_ref_topology(v.ctypes.data_as(_f32p), len(v), f.ctypes.data_as(_i32p), len(f), out.ctypes.data_as(_i32p),
                                ctypes.byref(levels), ctypes.byref(syncs))
    if n < 0:
        raise RuntimeError("reference hierarchy build did not terminate")
    return out[:n].copy(), int(levels.value), int(syncs.value)


def ref_last_fallback_rounds() -> int:
    """Rounds of the last ref_topology call in which some range took the order-statistic fallback (one sort each)."""
    lib().emul_ref_last_sorts.restype = ctypes.c_int
    return int(lib().emul_ref_last_sorts())
