"""Python mirror of ``lagrange::winding::FastWindingNumber`` over the C-ABI (include/wn_b200.h).

Reference surface (modules/winding/include/lagrange/winding/FastWindingNumber.h:24-108):
    FastWindingNumber(mesh); is_inside(pos) -> bool; solid_angle(pos) -> float
kept with the same names, argument meaning and error behaviour (non-3D / non-triangle meshes raise, like
la_runtime_assert at modules/winding/src/FastWindingNumber.cpp:91-96), plus the batched overloads BASELINE.json asks
for: arrays of points and implicit cell-centred lattices (the mesh_to_volume call pattern,
modules/volume/src/mesh_to_volume.cpp:147-149,175-182).

Every call goes to the CUDA engine; there is no host evaluation path. Inputs may be numpy arrays (host; staged by the
library) or torch CUDA tensors (used in place, asynchronous on the current torch stream).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _capi
from .mesh import SurfaceMesh

__all__ = ["FastWindingNumber", "Error"]


class Error(RuntimeError):
    """Counterpart of lagrange::Error (modules/core/include/lagrange/utils/Error.h)."""


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _current_stream_ptr():
    try:
        import torch

        if torch.cuda.is_available():
            return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    except Exception:
        pass
    return ctypes.c_void_p(0)


class _Buf:
    """A float32/int32/uint8 buffer the C-ABI can read or write: numpy (host) or torch (host or device)."""

    def __init__(self, obj, dtype, shape=None, writable=False):
        self.is_torch = _is_torch(obj)
        if self.is_torch:
            import torch

            tdt = {np.float32: torch.float32, np.int32: torch.int32, np.uint8: torch.uint8}[dtype]
            t = obj
            if t.dtype != tdt or not t.is_contiguous():
                if writable:
                    raise Error("output tensor must be contiguous and of the right dtype")
                t = t.to(tdt).contiguous()
            self.obj = t
            self.ptr = ctypes.c_void_p(t.data_ptr())
            self.size = t.numel()
            self.device = t.device
        else:
            a = np.asarray(obj)
            if writable:
                if a.dtype != dtype or not a.flags.c_contiguous or not a.flags.writeable:
                    raise Error("output array must be C-contiguous, writable and of the right dtype")
            else:
                a = np.ascontiguousarray(a, dtype=dtype)
            self.obj = a
            self.ptr = ctypes.c_void_p(a.ctypes.data)
            self.size = a.size
            self.device = None


def _alloc_like(points_buf, n, dtype):
    """Output buffer on the same side as the input: torch CUDA tensor for device inputs, numpy otherwise."""
    if points_buf is not None and points_buf.is_torch and points_buf.device.type == "cuda":
        import torch

        tdt = {np.float32: torch.float32, np.uint8: torch.uint8}[dtype]
        return torch.empty(n, dtype=tdt, device=points_buf.device)
    return np.empty(n, dtype=dtype)


class FastWindingNumber:
    """Fast winding number computation for triangle soups (B200 engine).

    Parameters mirror the reference constructor (a triangle ``SurfaceMesh``); ``vertices, facets`` arrays are accepted
    too. Coordinates are converted to float32 and indices to int32 like the reference does
    (FastWindingNumber.cpp:40-52).

    Options (additions to the reference surface):
      accuracy_scale  beta of the far-field test |q-P|^2 > beta^2 R^2 (reference default 2)
      order           Taylor order 0/1/2 (reference default 2)
      topology        (n_nodes, width) int32 child table to import instead of building the LBVH (oracle-tree mode)
      leaf_size, morton_bits, radius_mode ('box_corner' | 'vertex'), approximate_single_triangles, device
      hierarchy       'lbvh' (Morton + Karras, fastest build), 'kd' (balanced object-median splits), 'kd_sah' (k-d with binned
                      SAH cuts) or 'reference' (the reference builder's 4-ary top-down SAH tree, reproduced on the GPU: results
                      match the reference algorithm to float rounding)
    """

    def __init__(self, mesh=None, facets=None, *, accuracy_scale=2.0, order=2, topology=None, leaf_size=1, morton_bits=63,
                 radius_mode="box_corner", approximate_single_triangles=None, keep_build_data=False, device=None, hierarchy="reference",
                 _handle=None):
        self._h = None
        self._lib = _capi.lib()
        if _handle is not None:
            self._h = _handle
            return
        if mesh is None:
            # default-constructed engine, like FastWindingNumber() (FastWindingNumber.h:45): querying it raises
            return
        if isinstance(mesh, SurfaceMesh):
            if mesh.get_dimension() != 3:
                raise Error("Fast winding number engine only supports 3D meshes")
            if not mesh.is_triangle_mesh():
                raise Error("Fast winding number engine only supports triangle meshes")
            vertices, facets = mesh.vertices, mesh.facets
        else:
            vertices = mesh
            if facets is None:
                raise Error("FastWindingNumber(vertices, facets): facets missing")
        v = _Buf(vertices, np.float32)
        f = _Buf(facets, np.int32)
        vshape = tuple(v.obj.shape)
        fshape = tuple(f.obj.shape)
        if len(vshape) != 2 or vshape[1] != 3:
            raise Error("Fast winding number engine only supports 3D meshes")
        if len(fshape) != 2 or fshape[1] != 3:
            raise Error("Fast winding number engine only supports triangle meshes")
        opt = _capi.wn_options()
        _capi.check(self._lib.wn_options_init(ctypes.byref(opt)))
        opt.accuracy_scale = float(accuracy_scale)
        opt.order = int(order)
        opt.leaf_size = int(leaf_size)
        opt.morton_bits = int(morton_bits)
        opt.radius_mode = {"box_corner": _capi.WN_RADIUS_BOX_CORNER, "vertex": _capi.WN_RADIUS_VERTEX}[radius_mode]
        opt.keep_build_data = 1 if keep_build_data else 0
        if hierarchy not in _capi.WN_HIERARCHY:
            raise Error("hierarchy must be 'lbvh', 'kd', 'kd_sah' or 'reference'")
        opt.hierarchy = _capi.WN_HIERARCHY[hierarchy]
        opt.device = -1 if device is None else int(device)
        if approximate_single_triangles is None:
            approximate_single_triangles = topology is not None
        opt.approximate_single_triangles = 1 if approximate_single_triangles else 0
        h = ctypes.c_void_p()
        if topology is None:
            st = self._lib.wn_create(v.ptr, vshape[0], f.ptr, fshape[0], ctypes.byref(opt), ctypes.byref(h))
        else:
            topo = np.ascontiguousarray(topology, dtype=np.int32)
            if topo.ndim != 2:
                raise Error("topology must be a (n_nodes, width) child table")
            st = self._lib.wn_create_from_topology(v.ptr, vshape[0], f.ptr, fshape[0], ctypes.c_void_p(topo.ctypes.data), topo.shape[0],
                                                   topo.shape[1], ctypes.byref(opt), ctypes.byref(h))
        if st != _capi.WN_OK:
            raise Error(self._lib.wn_last_error().decode())
        self._h = h

    # -- lifetime ------------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.wn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _handle(self):
        if not self._h:
            raise Error("FastWindingNumber: engine is empty (default constructed or closed)")
        return self._h

    def _check(self, st):
        if st != _capi.WN_OK:
            raise Error(self._lib.wn_last_error().decode())

    @property
    def info(self) -> dict:
        i = _capi.wn_info()
        self._check(self._lib.wn_get_info(self._handle(), ctypes.byref(i)))
        return {k: getattr(i, k) for k, _ in _capi.wn_info._fields_ if k != "struct_size"}

    # -- the reference surface, single point or batched -------------------------------------------------------------
    def _points(self, pos):
        if _is_torch(pos):
            buf = _Buf(pos, np.float32)
            shape = tuple(buf.obj.shape)
        else:
            a = np.ascontiguousarray(pos, dtype=np.float32)
            buf = _Buf(a, np.float32)
            shape = a.shape
        single = len(shape) == 1
        if shape[-1] != 3 or len(shape) > 2:
            raise Error("query positions must have shape (3,) or (n, 3)")
        n = 1 if single else shape[0]
        return buf, n, single

    @staticmethod
    def _flags(presorted=False, tiling=True):
        return (_capi.WN_QUERY_PRESORTED if presorted else 0) | (0 if tiling else _capi.WN_QUERY_NO_TILING)

    def solid_angle(self, pos, accuracy_scale=None, presorted=False, out=None, tiling=True):
        """Solid angle at the query point(s). ``pos``: (3,) -> float, or (n,3) -> float32 array (numpy or torch CUDA)."""
        buf, n, single = self._points(pos)
        res = _alloc_like(buf, n, np.float32) if out is None else out
        ob = _Buf(res, np.float32, writable=True)
        flags = self._flags(presorted, tiling)
        self._check(self._lib.wn_solid_angle(self._handle(), buf.ptr, n, float(accuracy_scale or 0.0), flags, ob.ptr, _current_stream_ptr()))
        return float(res[0]) if single else res

    def is_inside(self, pos, accuracy_scale=None, presorted=False, out=None, tiling=True, bits=False):
        """True iff (double)solid_angle / (4 pi) > 0.5 (FastWindingNumber.cpp:66). (3,) -> bool, (n,3) -> uint8 array.
        ``bits=True``: the result is a bit array of (n + 7) // 8 bytes (``np.unpackbits(r, bitorder='little')[:n]``)."""
        buf, n, single = self._points(pos)
        bits = bits and not single
        res = _alloc_like(buf, (n + 7) // 8 if bits else n, np.uint8) if out is None else out
        ob = _Buf(res, np.uint8, writable=True)
        flags = self._flags(presorted, tiling) | (_capi.WN_QUERY_OUT_BITS if bits else 0)
        self._check(self._lib.wn_is_inside(self._handle(), buf.ptr, n, float(accuracy_scale or 0.0), flags, ob.ptr, _current_stream_ptr()))
        return bool(res[0]) if single else res

    # -- implicit lattice (mesh_to_volume pattern) --------------------------------------------------------------------
    @staticmethod
    def _grid_args(origin, spacing, dims, z_range):
        o = (ctypes.c_float * 3)(*[float(x) for x in origin])
        s = (ctypes.c_float * 3)(*[float(x) for x in spacing])
        d = (ctypes.c_int64 * 3)(*[int(x) for x in dims])
        z0, z1 = (0, int(dims[2])) if z_range is None else (int(z_range[0]), int(z_range[1]))
        n = int(dims[0]) * int(dims[1]) * max(0, z1 - z0)
        return o, s, d, z0, z1, n

    def query_grid(self, origin, spacing, dims, z_range=None, accuracy_scale=None, want_omega=False, want_inside=True, device_output=False,
                   out_omega=None, out_inside=None, tiling=True, layers=None, bits=False, shard=None):
        """Evaluate the cell-centred lattice p = origin + spacing*(ijk+0.5), x fastest; returns (omega, inside) (None if not wanted).

        ``device_output`` allocates torch CUDA outputs (results stay in HBM); otherwise numpy (copied to the host).
        ``layers=(first, step)``: strided multi-GPU sharding -- only the tile layers (8 z-planes each) first, first+step, ...
        are evaluated and returned compactly in that order (see ``strided_layer_planes``).
        ``shard=(rank, world)``: diagonal multi-GPU sharding (wn_query_grid_sharded): this rank's units -- one y-part of every c-th tile
        layer, see ``shard_layout`` -- returned compactly in layer order.
        ``bits=True``: ``inside`` is a bit array, point i -> bit (i & 7) of byte i >> 3 ((n + 7) // 8 bytes; WN_QUERY_OUT_BITS):
        an eighth of the device-to-host traffic for host outputs."""
        o, s, d, z0, z1, n = self._grid_args(origin, spacing, dims, z_range)
        if layers is not None:
            n = int(dims[0]) * int(dims[1]) * len(self.strided_layer_planes(int(dims[2]), *layers))
        if shard is not None:
            n = self.shard_layout(dims, *shard)["n_points"]
        def mk(dtype, given, want, count):
            if given is not None:
                return given
            if not want:
                return None
            if device_output:
                import torch

                return torch.empty(count, dtype={np.float32: torch.float32, np.uint8: torch.uint8}[dtype], device="cuda")
            return np.empty(count, dtype=dtype)
        om = mk(np.float32, out_omega, want_omega, n)
        ins = mk(np.uint8, out_inside, want_inside, (n + 7) // 8 if bits else n)
        pom = _Buf(om, np.float32, writable=True).ptr if om is not None else None
        pin = _Buf(ins, np.uint8, writable=True).ptr if ins is not None else None
        flags = self._flags(False, tiling) | (_capi.WN_QUERY_OUT_BITS if (bits and ins is not None) else 0)
        if shard is not None:
            self._check(self._lib.wn_query_grid_sharded(self._handle(), o, s, d, int(shard[0]), int(shard[1]), float(accuracy_scale or 0.0), flags, pom,
                                                        pin, _current_stream_ptr()))
        elif layers is not None:
            self._check(self._lib.wn_query_grid_strided(self._handle(), o, s, d, int(layers[0]), int(layers[1]), float(accuracy_scale or 0.0),
                                                        flags, pom, pin, _current_stream_ptr()))
        else:
            self._check(self._lib.wn_query_grid(self._handle(), o, s, d, z0, z1, float(accuracy_scale or 0.0), flags, pom,
                                                pin, _current_stream_ptr()))
        return om, ins

    @staticmethod
    def shard_layout(dims, rank, world) -> dict:
        """Layout of ``query_grid(shard=(rank, world))``'s output: {parts_y, layer_step, part_rows, n_units, n_points, units}, with
        units = [(z0, z1, y0, y1), ...] in output order (each unit holds (z1 - z0) * (y1 - y0) * nx values, x fastest)."""
        lib = _capi.lib()
        d = (ctypes.c_int64 * 3)(*[int(x) for x in dims])
        q, c = ctypes.c_int32(), ctypes.c_int32()
        rows, nu, npts = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        _capi.check(lib.wn_grid_shard_layout(d, int(rank), int(world), ctypes.byref(q), ctypes.byref(c), ctypes.byref(rows), ctypes.byref(nu),
                                             ctypes.byref(npts)))
        Q, C, R, nz = q.value, c.value, rows.value, int(dims[2])
        units = []
        for k in range(nu.value):
            lz = rank % C + k * C
            part = ((rank - lz) // C) % Q
            units.append((8 * lz, min(nz, 8 * lz + 8), part * R, (part + 1) * R))
        return {"parts_y": Q, "layer_step": C, "part_rows": R, "n_units": nu.value, "n_points": npts.value, "units": units}

    @staticmethod
    def strided_layer_planes(nz, first, step):
        """z indices, in output order, of the planes that ``query_grid(layers=(first, step))`` evaluates."""
        out = []
        for layer in range(int(first), (nz + 7) // 8, int(step)):
            out.extend(range(layer * 8, min(nz, layer * 8 + 8)))
        return out

    def sdf_grid(self, origin, spacing, dims, band, accuracy_scale=None, signed=True, device_output=False, out=None, tiling=True):
        """Narrow-band distance to the mesh at the lattice's cell centres, negative where ``is_inside`` holds, clamped to
        +-band (world units). Returns (sdf [nz, ny, nx] float32, number of cells with distance < band). wn_sdf_grid."""
        o, s, d, _, _, n = self._grid_args(origin, spacing, dims, None)
        if out is None:
            if device_output:
                import torch

                out = torch.empty(n, dtype=torch.float32, device="cuda")
            else:
                out = np.empty(n, dtype=np.float32)
        flags = self._flags(False, tiling) | (0 if signed else 4)
        active = ctypes.c_int64()
        self._check(self._lib.wn_sdf_grid(self._handle(), o, s, d, float(band), float(accuracy_scale or 0.0), flags,
                                          _Buf(out, np.float32, writable=True).ptr, ctypes.byref(active), _current_stream_ptr()))
        return out.reshape(int(dims[2]), int(dims[1]), int(dims[0])), int(active.value)

    def sdf_grid_sparse(self, origin, spacing, dims, band, accuracy_scale=None, signed=True, tiling=True, want_inside_bits=False):
        """Narrow band only: (linear indices int64 [m], signed distances float32 [m], inside bits or None). The cells with
        |distance| < band, ordered by linear index (z*ny + y)*nx + x: what an OpenVDB grid keeps active. wn_sdf_grid_sparse."""
        o, s, d, _, _, n = self._grid_args(origin, spacing, dims, None)
        flags = self._flags(False, tiling) | (0 if signed else 4)
        active = ctypes.c_int64()
        self._check(self._lib.wn_sdf_grid_sparse(self._handle(), o, s, d, float(band), float(accuracy_scale or 0.0), flags, 0, None, None, None,
                                                 ctypes.byref(active), _current_stream_ptr()))
        m = int(active.value)
        idx = np.empty(m, dtype=np.int64)
        val = np.empty(m, dtype=np.float32)
        bits = np.empty((n + 7) // 8, dtype=np.uint8) if want_inside_bits else None
        self._check(self._lib.wn_sdf_grid_sparse(self._handle(), o, s, d, float(band), float(accuracy_scale or 0.0), flags, m,
                                                 ctypes.c_void_p(idx.ctypes.data) if m else None, ctypes.c_void_p(val.ctypes.data) if m else None,
                                                 ctypes.c_void_p(bits.ctypes.data) if bits is not None else None, ctypes.byref(active),
                                                 _current_stream_ptr()))
        return idx, val, bits

    def closest_point(self, pos, max_distance=None, presorted=False):
        """Closest point on the mesh: (squared distance, triangle id, closest point) per query, like
        TriangleAABBTree::get_closest_point (modules/bvh/include/lagrange/bvh/TriangleAABBTree.h:84-88). ``pos``: (3,) or (n,3),
        numpy or torch CUDA. ``max_distance`` bounds the search (farther points report it, triangle -1)."""
        buf, n, single = self._points(pos)
        if buf.is_torch and buf.device.type == "cuda":
            import torch

            sq = torch.empty(n, dtype=torch.float32, device=buf.device)
            tri = torch.empty(n, dtype=torch.int32, device=buf.device)
            xyz = torch.empty((n, 3), dtype=torch.float32, device=buf.device)
            ptrs = [ctypes.c_void_p(t.data_ptr()) for t in (sq, tri, xyz)]
        else:
            sq, tri, xyz = np.empty(n, dtype=np.float32), np.empty(n, dtype=np.int32), np.empty((n, 3), dtype=np.float32)
            ptrs = [ctypes.c_void_p(t.ctypes.data) for t in (sq, tri, xyz)]
        self._check(self._lib.wn_closest_point(self._handle(), buf.ptr, n, float(max_distance or 0.0), self._flags(presorted, True) & 1, ptrs[0], ptrs[1],
                                               ptrs[2], _current_stream_ptr()))
        if single:
            return float(sq[0]), int(tri[0]), xyz[0]
        return sq, tri, xyz

    def is_inside_grid(self, origin, spacing, dims, **kw):
        return self.query_grid(origin, spacing, dims, want_inside=True, want_omega=False, **kw)[1]

    def solid_angle_grid(self, origin, spacing, dims, **kw):
        return self.query_grid(origin, spacing, dims, want_inside=False, want_omega=True, **kw)[0]

    # -- exact brute force mode ------------------------------------------------------------------------------------------
    def exact_solid_angle(self, pos, out=None):
        buf, n, single = self._points(pos)
        res = _alloc_like(buf, n, np.float32) if out is None else out
        ob = _Buf(res, np.float32, writable=True)
        self._check(self._lib.wn_exact(self._handle(), buf.ptr, n, ob.ptr, None, _current_stream_ptr()))
        return float(res[0]) if single else res

    def exact_is_inside(self, pos, out=None):
        buf, n, single = self._points(pos)
        res = _alloc_like(buf, n, np.uint8) if out is None else out
        ob = _Buf(res, np.uint8, writable=True)
        self._check(self._lib.wn_exact(self._handle(), buf.ptr, n, None, ob.ptr, _current_stream_ptr()))
        return bool(res[0]) if single else res

    def exact_grid(self, origin, spacing, dims, z_range=None, want_omega=True, want_inside=False):
        o, s, d, z0, z1, n = self._grid_args(origin, spacing, dims, z_range)
        om = np.empty(n, dtype=np.float32) if want_omega else None
        ins = np.empty(n, dtype=np.uint8) if want_inside else None
        pom = ctypes.c_void_p(om.ctypes.data) if om is not None else None
        pin = ctypes.c_void_p(ins.ctypes.data) if ins is not None else None
        self._check(self._lib.wn_exact_grid(self._handle(), o, s, d, z0, z1, pom, pin, _current_stream_ptr()))
        return om, ins

    # -- counters ------------------------------------------------------------------------------------------------------
    def query_stats(self, pos, accuracy_scale=None, presorted=False, tiling=False) -> dict:
        """Executed-work counters. tiling=False (default): the reference algorithm's per-point counts."""
        buf, n, _ = self._points(pos)
        st = _capi.wn_query_stats()
        flags = self._flags(presorted, tiling)
        self._check(self._lib.wn_query_stats_points(self._handle(), buf.ptr, n, float(accuracy_scale or 0.0), flags, ctypes.byref(st),
                                                    _current_stream_ptr()))
        return self._stats_dict(st, n)

    def query_stats_grid(self, origin, spacing, dims, z_range=None, accuracy_scale=None, tiling=False) -> dict:
        o, s, d, z0, z1, n = self._grid_args(origin, spacing, dims, z_range)
        st = _capi.wn_query_stats()
        self._check(self._lib.wn_query_stats_grid(self._handle(), o, s, d, z0, z1, float(accuracy_scale or 0.0), self._flags(False, tiling),
                                                  ctypes.byref(st), _current_stream_ptr()))
        return self._stats_dict(st, n)

    @staticmethod
    def _stats_dict(st, n):
        T, A, E, V = st.node_tests, st.far_field_evals, st.exact_triangles, st.lane_slots
        return {"queries": n, "node_tests": T, "far_field_evals": A, "exact_triangles": E, "lane_slots": V,
                # 10 flop per test, 83 more per accepted far-field evaluation, 75 per exact triangle
                "algorithmic_flops": 10 * T + 83 * A + 75 * E}

    # -- replication across GPUs -----------------------------------------------------------------------------------------
    def packed_size(self) -> int:
        n = ctypes.c_int64()
        self._check(self._lib.wn_tree_packed_size(self._handle(), ctypes.byref(n)))
        return int(n.value)

    def pack(self, out=None):
        """Serialise the packed tree into a uint8 buffer (torch CUDA tensor if given/available, else numpy)."""
        n = self.packed_size()
        if out is None:
            out = np.empty(n, dtype=np.uint8)
        ob = _Buf(out, np.uint8, writable=True)
        self._check(self._lib.wn_tree_pack(self._handle(), ob.ptr, ob.size, _current_stream_ptr()))
        return out

    @classmethod
    def from_packed(cls, buf, accuracy_scale=None, device=None):
        lib = _capi.lib()
        b = _Buf(buf, np.uint8)
        h = ctypes.c_void_p()
        if b.is_torch and b.device.type == "cuda":
            import torch

            torch.cuda.current_stream().synchronize()  # the adopt copy runs on the default stream
        opt = None
        if accuracy_scale is not None or device is not None:
            opt = _capi.wn_options()
            _capi.check(lib.wn_options_init(ctypes.byref(opt)))
            # <= 0 = keep the accuracy scale stored with the tree (a replica answers like the engine it was packed from)
            opt.accuracy_scale = float(accuracy_scale) if accuracy_scale is not None else 0.0
            if device is not None:
                opt.device = int(device)
        st = lib.wn_create_from_packed(b.ptr, b.size, ctypes.byref(opt) if opt is not None else None, ctypes.byref(h))
        if st != _capi.WN_OK:
            raise Error(lib.wn_last_error().decode())
        return cls(_handle=h)

    def replicate(self, devices):
        """Copies of this engine on other GPUs of the same process (wn_replicate: device-to-device copies of the packed tree)."""
        devices = [int(d) for d in devices]
        arr = (ctypes.c_int32 * len(devices))(*devices)
        out = (ctypes.c_void_p * len(devices))()
        self._check(self._lib.wn_replicate(self._handle(), arr, len(devices), out))
        return [FastWindingNumber(_handle=ctypes.c_void_p(h)) for h in out]

    @staticmethod
    def query_grid_multi(engines, origin, spacing, dims, accuracy_scale=None, want_omega=False, want_inside=True, tiling=True, bits=False):
        """One lattice over several engines holding the same tree (one per GPU, single process): engine i evaluates the tile layers
        i, i+N, ... from its own host thread; results are gathered in lattice order on the host (wn_query_grid_multi)."""
        lib = _capi.lib()
        o, s, d, _, _, n = FastWindingNumber._grid_args(origin, spacing, dims, None)
        om = np.empty(n, dtype=np.float32) if want_omega else None
        ins = np.empty((n + 7) // 8 if bits else n, dtype=np.uint8) if want_inside else None
        hs = (ctypes.c_void_p * len(engines))(*[e._handle() for e in engines])
        flags = FastWindingNumber._flags(False, tiling) | (_capi.WN_QUERY_OUT_BITS if (bits and ins is not None) else 0)
        st = lib.wn_query_grid_multi(hs, len(engines), o, s, d, float(accuracy_scale or 0.0), flags,
                                     ctypes.c_void_p(om.ctypes.data) if om is not None else None,
                                     ctypes.c_void_p(ins.ctypes.data) if ins is not None else None)
        if st != _capi.WN_OK:
            raise Error(lib.wn_last_error().decode())
        return om, ins

    # -- parity hooks ------------------------------------------------------------------------------------------------------
    def debug_node_moments(self, first=0, count=None) -> np.ndarray:
        if count is None:
            count = self.info["num_tree_nodes"] - first
        out = np.empty((count, 23), dtype=np.float32)
        self._check(self._lib.wn_debug_node_moments(self._handle(), first, count, ctypes.c_void_p(out.ctypes.data)))
        return out

    def debug_last_plan(self) -> np.ndarray:
        """[tiles, 4] int32: conditional records, direct records, exact triangles, flags of the last tiled batch."""
        n = ctypes.c_int64()
        self._check(self._lib.wn_debug_last_plan(self._handle(), None, 0, ctypes.byref(n)))
        out = np.empty((n.value, 4), dtype=np.int32)
        if n.value:
            self._check(self._lib.wn_debug_last_plan(self._handle(), ctypes.c_void_p(out.ctypes.data), n.value, ctypes.byref(n)))
        return out

    def debug_topology(self) -> np.ndarray:
        n = ctypes.c_int64()
        self._check(self._lib.wn_debug_topology(self._handle(), None, 0, ctypes.byref(n)))
        w = self.info["width"]
        out = np.empty((n.value, w), dtype=np.int32)
        self._check(self._lib.wn_debug_topology(self._handle(), ctypes.c_void_p(out.ctypes.data), n.value, ctypes.byref(n)))
        return out
