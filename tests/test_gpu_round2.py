"""GPU tier (-m gpu), round 2 additions: bit-packed outputs, packed-tree adoption, corrupted blobs, lattice bounds.

Everything goes through the C-ABI (ctypes). Parity here is against this repo's restatement of the reference algorithm
(oracle/, "parity unpinned": no reference-held vectors exist for this path), never against the upstream binary.
"""
import ctypes

import numpy as np
import pytest

from conftest import small_config

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lb():
    import torch

    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    import lagrange_b200

    return lagrange_b200


def _unpack(bits, n):
    return np.unpackbits(np.asarray(bits, dtype=np.uint8), bitorder="little")[:n]


@pytest.mark.parametrize("tiling", [True, False])
@pytest.mark.parametrize("dims", [(64, 64, 64), (37, 21, 19), (5, 5, 11), (130, 9, 33)])
def test_bit_packed_lattice_output_equals_bytes(lb, prim, dims, tiling):
    """WN_QUERY_OUT_BITS: bit i of the packed output == byte i of the plain output, host and device destinations,
    whole lattice, z-slab and strided layers (ragged sizes: n not a multiple of 8, partial last layer)."""
    import torch

    V, F = prim.generate_torus(5.0, 1.0, 64, 32)
    eng = lb.FastWindingNumber(V, F, hierarchy="kd_sah", leaf_size=4)
    lo, hi = prim.mesh_bbox(V)
    origin = (lo - 0.3).astype(np.float32)
    spacing = ((hi - lo + 0.6) / np.array(dims)).astype(np.float32)
    d = np.array(dims, dtype=np.int64)
    n = int(np.prod(d))
    _, ref = eng.query_grid(origin, spacing, d, tiling=tiling)
    assert 0 < ref.sum() < n
    _, bits = eng.query_grid(origin, spacing, d, tiling=tiling, bits=True)
    assert bits.shape == ((n + 7) // 8,)
    assert np.array_equal(_unpack(bits, n), ref)
    # device destination
    out = torch.zeros((n + 7) // 8, dtype=torch.uint8, device="cuda")
    eng.query_grid(origin, spacing, d, tiling=tiling, bits=True, out_inside=out)
    assert np.array_equal(_unpack(out.cpu().numpy(), n), ref)
    # z-slab
    z0, z1 = 3, int(d[2]) - 2
    _, sb = eng.query_grid(origin, spacing, d, z_range=(z0, z1), tiling=tiling, bits=True)
    per = int(d[0] * d[1])
    assert np.array_equal(_unpack(sb, per * (z1 - z0)), ref[z0 * per:z1 * per])
    # strided layers (multi-GPU sharding): compact output of layers first, first + step, ...
    for first, step in ((0, 2), (1, 2), (2, 3)):
        planes = eng.strided_layer_planes(int(d[2]), first, step)
        if not planes:
            continue
        _, lb_bits = eng.query_grid(origin, spacing, d, layers=(first, step), tiling=tiling, bits=True)
        want = np.concatenate([ref[z * per:(z + 1) * per] for z in planes])
        assert np.array_equal(_unpack(lb_bits, len(want)), want)


def test_bit_packed_point_output(lb, prim):
    V, F, P, _ = small_config(prim, 5)
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    for n in (1, 7, 8, 9, 1000, 4099, len(P)):
        ref = eng.is_inside(P[:n])
        bits = eng.is_inside(P[:n], bits=True)
        assert bits.shape == ((n + 7) // 8,)
        assert np.array_equal(_unpack(bits, n), ref)


def test_bit_packed_large_host_output_is_pipelined_and_equal(lb, prim):
    """A lattice large enough for the tiled path to split into batches (host output => copies overlap the next batch)."""
    V, F = prim.generate_subdivided_sphere("icosahedron", 5)
    eng = lb.FastWindingNumber(V, F, hierarchy="kd_sah", leaf_size=4)
    n = 192
    origin = np.full(3, -1.1, np.float32)
    spacing = np.full(3, 2.2 / n, np.float32)
    d = np.array([n, n, n], dtype=np.int64)
    _, ref = eng.query_grid(origin, spacing, d)
    _, bits = eng.query_grid(origin, spacing, d, bits=True)
    assert np.array_equal(_unpack(bits, n ** 3), ref)


def test_from_packed_keeps_the_accuracy_scale_of_the_tree(lb, prim):
    """a replica adopted with options (device=...) must keep the packed engine's beta, so that
    beta <= 0 queries answer identically on the source and on every replica."""
    V, F, P, _ = small_config(prim, 1)
    src = lb.FastWindingNumber(V, F, hierarchy="lbvh", accuracy_scale=3.0)
    blob = src.pack()
    rep = lb.FastWindingNumber.from_packed(blob, device=0)  # options given, accuracy_scale not
    assert rep.info["accuracy_scale"] == pytest.approx(3.0)
    assert np.array_equal(rep.solid_angle(P, tiling=False), src.solid_angle(P, tiling=False))
    assert not np.array_equal(src.solid_angle(P, tiling=False), src.solid_angle(P, accuracy_scale=2.0, tiling=False))
    rep2 = lb.FastWindingNumber.from_packed(blob, accuracy_scale=2.0)  # explicit override still honoured
    assert rep2.info["accuracy_scale"] == pytest.approx(2.0)
    assert np.array_equal(rep2.solid_angle(P, tiling=False), src.solid_angle(P, accuracy_scale=2.0, tiling=False))
    rep3 = lb.FastWindingNumber.from_packed(blob)
    assert rep3.info["accuracy_scale"] == pytest.approx(3.0)


def test_corrupted_packed_tree_is_rejected(lb, prim):
    """links / child indices / leaf ranges inside an adopted blob are validated on the device."""
    V, F = prim.generate_torus(5.0, 1.0, 24, 12)
    src = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    blob = src.pack()
    hdr = np.frombuffer(blob[:128].tobytes(), dtype=np.int64)
    n_entries, off_hot, off_kids = int(hdr[2]), int(hdr[4]), int(hdr[6])
    assert n_entries == src.info["num_entries"]
    # (a) a skip link pointing far outside the array
    bad = blob.copy()
    link = bad[off_hot:off_hot + 32 * n_entries].view(np.int32).reshape(n_entries, 8)
    internal = np.nonzero(link[:, 3] >= 0)[0]  # R2 sign bit clear = internal entry
    link[internal[1], 7] = n_entries + 12345
    with pytest.raises(lb.Error):
        lb.FastWindingNumber.from_packed(bad)
    # (b) a child index out of range
    bad = blob.copy()
    kids = bad[off_kids:off_kids + 16 * n_entries].view(np.int32).reshape(n_entries, 4)
    kids[0, 1] = n_entries + 7
    with pytest.raises(lb.Error):
        lb.FastWindingNumber.from_packed(bad)
    # (c) a leaf whose triangle range runs past the triangle array
    bad = blob.copy()
    link = bad[off_hot:off_hot + 32 * n_entries].view(np.int32).reshape(n_entries, 8)
    leaves = np.nonzero(link[:, 3] < 0)[0]
    link[leaves[-1], 7] = (len(F) << 4) | 3
    with pytest.raises(lb.Error):
        lb.FastWindingNumber.from_packed(bad)
    # (d) negative counts in the header
    bad = blob.copy()
    bad[:128].view(np.int64)[3] = -5
    with pytest.raises(lb.Error):
        lb.FastWindingNumber.from_packed(bad)
    # the intact blob still loads
    assert lb.FastWindingNumber.from_packed(blob).info["num_entries"] == n_entries


def test_oversized_lattices_are_rejected_not_wrapped(lb, prim):
    """dims up to 2^24 each used to overflow the point count."""
    V, F = prim.generate_torus(5.0, 1.0, 24, 12)
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    from lagrange_b200 import _capi

    L = _capi.lib()
    o = (ctypes.c_float * 3)(0, 0, 0)
    s = (ctypes.c_float * 3)(1, 1, 1)
    d = (ctypes.c_int64 * 3)(1 << 24, 1 << 24, 1 << 24)
    out = np.zeros(16, dtype=np.uint8)
    st = L.wn_query_grid(eng._handle(), o, s, d, 0, 1 << 24, 0.0, 0, None, ctypes.c_void_p(out.ctypes.data), None)
    assert st == 4  # WN_ERR_UNSUPPORTED
    assert b"too large" in L.wn_last_error()


# ---- closest point on the mesh (TriangleAABBTree::get_closest_point) ---------------------------------------
def _tri_closest64(p, a, b, c):
    """Closest point of triangle (a, b, c) to p in float64 (Ericson 5.1.5), vectorised over rows."""
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = (ab * ap).sum(1), (ac * ap).sum(1)
    bp = p - b
    d3, d4 = (ab * bp).sum(1), (ac * bp).sum(1)
    cp = p - c
    d5, d6 = (ab * cp).sum(1), (ac * cp).sum(1)
    vc, vb, va = d1 * d4 - d3 * d2, d5 * d2 - d1 * d6, d3 * d6 - d5 * d4
    out = np.empty_like(p)
    done = np.zeros(len(p), bool)

    def put(mask, val):
        m = mask & ~done
        out[m] = val[m]
        done[m] = True

    put((d1 <= 0) & (d2 <= 0), a)
    put((d3 >= 0) & (d4 <= d3), b)
    with np.errstate(divide="ignore", invalid="ignore"):
        put((vc <= 0) & (d1 >= 0) & (d3 <= 0), a + (d1 / (d1 - d3))[:, None] * ab)
        put((d6 >= 0) & (d5 <= d6), c)
        put((vb <= 0) & (d2 >= 0) & (d6 <= 0), a + (d2 / (d2 - d6))[:, None] * ac)
        put((va <= 0) & (d4 - d3 >= 0) & (d5 - d6 >= 0), b + ((d4 - d3) / ((d4 - d3) + (d5 - d6)))[:, None] * (c - b))
        den = 1.0 / (va + vb + vc)
        put(np.ones(len(p), bool), a + ab * (vb * den)[:, None] + ac * (vc * den)[:, None])
    return out


@pytest.mark.parametrize("hierarchy,leaf", [("reference", 1), ("kd_sah", 4), ("lbvh", 1)])
def test_closest_point_matches_brute_force(lb, oracle_mod, prim, hierarchy, leaf):
    """wn_closest_point vs the double-precision brute force (oracle.distance64): the squared distance agrees, the returned
    triangle attains it, and the returned point is that triangle's closest point (ties between triangles sharing the closest
    edge or vertex are legitimate, so the id is checked through the distance, like the reference's own contract)."""
    import torch

    V, F = prim.config_mesh(3, small=True)  # open soup with duplicates and flips
    eng = lb.FastWindingNumber(V, F, hierarchy=hierarchy, leaf_size=leaf)
    lo, hi = prim.mesh_bbox(V)
    rng = np.random.Generator(np.random.PCG64(11))
    P = np.concatenate([prim.uniform_points_in_bbox(lo, hi, 20000, seed=5), prim.near_surface_points(V, F, 20000, sigma_rel=5e-3, seed=6),
                        V[rng.integers(0, len(V), 500)],  # on the mesh: distance 0
                        (rng.random((200, 3)) * 200 - 100).astype(np.float32)]).astype(np.float32)  # far away
    sq, tri, xyz = eng.closest_point(P)
    d_ref = oracle_mod.distance64(V, F, P)
    scale = float(np.linalg.norm(hi - lo))
    assert np.abs(np.sqrt(sq.astype(np.float64)) - d_ref).max() < 2e-6 * max(scale, float(d_ref.max()))
    assert tri.min() >= 0 and tri.max() < len(F)
    Vd = V.astype(np.float64)
    a, b, c = Vd[F[tri, 0]], Vd[F[tri, 1]], Vd[F[tri, 2]]
    cp = _tri_closest64(P.astype(np.float64), a, b, c)
    d_tri = np.linalg.norm(P - cp, axis=1)
    assert np.abs(d_tri - d_ref).max() < 2e-6 * max(scale, float(d_ref.max()))  # the triangle attains the minimum
    assert np.abs(xyz - cp).max() < 1e-5 * scale  # and the point is the closest point of that triangle
    # device pointers, same answers; single point overload
    sq_d, tri_d, xyz_d = eng.closest_point(torch.from_numpy(P).cuda())
    assert np.array_equal(sq_d.cpu().numpy(), sq) and np.array_equal(tri_d.cpu().numpy(), tri) and np.array_equal(xyz_d.cpu().numpy(), xyz)
    s1, t1, x1 = eng.closest_point(P[7])
    assert s1 == sq[7] and np.array_equal(x1, xyz[7])
    # bounded search: points farther than the bound report it
    bound = float(np.median(d_ref))
    sq_b, tri_b, _ = eng.closest_point(P, max_distance=bound)
    near = d_ref < bound * (1 - 1e-5)
    far = d_ref > bound * (1 + 1e-5)
    assert np.array_equal(sq_b[near], sq[near]) and np.all(tri_b[far] == -1) and np.allclose(sq_b[far], bound * bound, rtol=1e-6)


def test_sparse_narrow_band_equals_the_dense_block(lb, prim):
    V, F = prim.generate_subdivided_sphere("icosahedron", 4)
    eng = lb.FastWindingNumber(V, F)
    for dims in ((48, 48, 48), (37, 21, 19)):
        d = np.array(dims, dtype=np.int64)
        o = np.full(3, -1.2, np.float32)
        s = (2.4 / d).astype(np.float32)
        band = 3.0 * float(s.max())
        dense, n_active = eng.sdf_grid(o, s, d, band)
        idx, val, bits = eng.sdf_grid_sparse(o, s, d, band, want_inside_bits=True)
        flat = dense.reshape(-1)
        want = np.nonzero(np.abs(flat) < band)[0]
        assert len(idx) == n_active == len(want)
        assert np.array_equal(idx, want) and np.array_equal(val, flat[want])
        assert np.array_equal(np.unpackbits(bits, bitorder="little")[: flat.size].astype(bool), flat < 0)
        assert 0 < len(idx) < flat.size


def test_exact_mode_grouped_fast_path_keeps_the_zero_rules_and_accuracy(lb, oracle_mod, prim):
    """K7 folds groups of 8 small-angle triangles into one complex product + one atan2; a group with a triangle the query is
    close to (or on) is redone term by term with the reference formulation (A.1: a vertex on the query or a zero numerator
    contribute 0). Queries ON vertices, IN triangle planes (inside and outside the triangle), very close to the surface and
    far away all have to agree with the double-precision brute force."""
    V, F = prim.generate_torus(5.0, 1.0, 64, 32)  # 4096 triangles: 16 shared-memory tiles of 256
    rng = np.random.Generator(np.random.PCG64(3))
    on_vertex = V[rng.integers(0, len(V), 300)]
    tri = F[rng.integers(0, len(F), 600)]
    w = rng.dirichlet(np.ones(3), size=600).astype(np.float32)
    in_plane_inside = (V[tri[:, 0]] * w[:, :1] + V[tri[:, 1]] * w[:, 1:2] + V[tri[:, 2]] * w[:, 2:3]).astype(np.float32)
    e1, e2 = V[tri[:, 1]] - V[tri[:, 0]], V[tri[:, 2]] - V[tri[:, 0]]
    in_plane_outside = (V[tri[:, 0]] + 3.0 * e1 + 2.0 * e2).astype(np.float32)
    near = prim.near_surface_points(V, F, 5000, sigma_rel=1e-4, seed=9)
    far = prim.uniform_points_in_bbox(*prim.mesh_bbox(V), 5000, inflate=2.0, seed=10)
    P = np.concatenate([on_vertex, in_plane_outside, near, far, in_plane_inside]).astype(np.float32)
    eng = lb.FastWindingNumber(V, F)
    got = eng.exact_solid_angle(P)
    want = oracle_mod.exact64(V, F, P)
    assert np.all(np.isfinite(got))
    # (a point INSIDE a triangle, in its plane, sits on the 4 pi jump of that triangle's solid angle: which side float rounding puts
    # it on is arbitrary in any implementation, so those are only required to be finite and within 4 pi of the reference)
    n_ok = len(P) - len(in_plane_inside)
    err = np.abs(got[:n_ok] - want[:n_ok])
    n_a, n_b = len(on_vertex) + len(in_plane_outside), len(on_vertex) + len(in_plane_outside) + len(near)
    assert err[:n_a].max() < 2e-5 * 4 * np.pi and err[n_b:].max() < 2e-5 * 4 * np.pi
    # 1e-4 of the bbox diagonal from the surface the float32 formulation itself is ill-conditioned (a triangle fills almost a half
    # space): the bar there is the float32 twin of the reference formulation on the CPU, and BASELINE's 1e-4 * 4 pi against double
    twin = oracle_mod.exact32(V, F, P[n_a:n_b])
    assert err[n_a:n_b].max() < 1e-4 * 4 * np.pi
    assert np.abs(got[n_a:n_b] - twin).max() < 1e-4 * 4 * np.pi
    d = np.abs(got[n_ok:] - want[n_ok:])
    assert np.all((d < 1e-4 * 4 * np.pi) | (np.abs(d - 4 * np.pi) < 1e-4 * 4 * np.pi) | (np.abs(d - 2 * np.pi) < 1e-4 * 4 * np.pi))
    # small batches take the warp-per-query kernel: same function, same answers
    small = eng.exact_solid_angle(P[:100])
    assert np.abs(small - got[:100]).max() < 1e-5 * 4 * np.pi


@pytest.mark.parametrize("dims", [(72, 40, 56), (100, 100, 100), (33, 65, 129)])
def test_hierarchical_planning_on_ragged_lattices(lb, prim, monkeypatch, dims):
    """The tiled path with planning blocks above the tiles (2x2x2 / 4x4x4 tiles; flat blocks for strided layers), forced on lattices
    whose tile counts are not multiples of the block sizes: same results as the per-point traversal up to the far-field
    interpolation (<= 3e-5 * 4 pi), with and without the block levels, whole lattice and every-3rd-layer sharding."""
    monkeypatch.setenv("WN_TILE", "1")
    V, F = prim.generate_torus(5.0, 1.0, 100, 50)
    eng = lb.FastWindingNumber(V, F)
    lo, hi = prim.mesh_bbox(V)
    d = np.array(dims, dtype=np.int64)
    o = (lo - 0.4).astype(np.float32)
    s = ((hi - lo + 0.8) / d).astype(np.float32)
    ref = eng.query_grid(o, s, d, want_omega=True, tiling=False)[0]
    tol = 3e-5 * 4 * np.pi
    per = int(d[0] * d[1])
    for levels in ("2", "0", "1", "3"):
        monkeypatch.setenv("WN_PLAN_LEVELS", levels)
        om = eng.query_grid(o, s, d, want_omega=True)[0]
        assert np.abs(om - ref).max() < tol, (levels, float(np.abs(om - ref).max()))
        for first, step in ((1, 3), (0, 2)):
            planes = eng.strided_layer_planes(int(d[2]), first, step)
            part = eng.query_grid(o, s, d, want_omega=True, layers=(first, step))[0]
            want = np.concatenate([ref[z * per:(z + 1) * per] for z in planes])
            assert np.abs(part - want).max() < tol, (levels, first, step)
    # the plan really ran on this lattice (otherwise this test checks nothing)
    st_t = eng.query_stats_grid(o, s, d, tiling=True)
    st_g = eng.query_stats_grid(o, s, d, tiling=False)
    assert st_t["node_tests"] < st_g["node_tests"]


@pytest.mark.parametrize("world", [2, 4, 8, 6])
@pytest.mark.parametrize("dims", [(64, 64, 64), (96, 64, 100), (40, 24, 37)])
def test_diagonal_sharding_reassembles_the_lattice(lb, prim, monkeypatch, dims, world):
    """wn_query_grid_sharded: rank r takes one y-part of every c-th tile layer; the ranks' compact outputs, laid out with
    distributed.gather_sharded, are the whole lattice (per-point path: bit-identical; tiled path: up to the far-field interpolation;
    bit-packed outputs: the same bits)."""
    from lagrange_b200.distributed import gather_sharded

    V, F = prim.generate_torus(5.0, 1.0, 64, 32)
    eng = lb.FastWindingNumber(V, F)
    lo, hi = prim.mesh_bbox(V)
    d = np.array(dims, dtype=np.int64)
    o = (lo - 0.3).astype(np.float32)
    s = ((hi - lo + 0.6) / d).astype(np.float32)
    om_ref, in_ref = eng.query_grid(o, s, d, want_omega=True, tiling=False)
    parts_om, parts_in = [], []
    for r in range(world):
        om, ins = eng.query_grid(o, s, d, want_omega=True, tiling=False, shard=(r, world))
        parts_om.append(om)
        parts_in.append(ins)
    assert np.array_equal(gather_sharded(d, world, parts_om).reshape(-1), om_ref)
    assert np.array_equal(gather_sharded(d, world, parts_in).reshape(-1), in_ref)
    monkeypatch.setenv("WN_TILE", "1")  # the tiled path (hierarchical planning with flat blocks inside a part)
    parts_om, parts_bits = [], []
    for r in range(world):
        parts_om.append(eng.query_grid(o, s, d, want_omega=True, shard=(r, world))[0])
        n_r = lb.FastWindingNumber.shard_layout(d, r, world)["n_points"]
        bits = eng.query_grid(o, s, d, shard=(r, world), bits=True)[1]
        byt = eng.query_grid(o, s, d, shard=(r, world))[1]
        assert np.array_equal(np.unpackbits(bits, bitorder="little")[:n_r], byt)
        parts_bits.append(byt)
    om_t = gather_sharded(d, world, parts_om).reshape(-1)
    assert np.abs(om_t - om_ref).max() < 3e-5 * 4 * np.pi
    in_t = gather_sharded(d, world, parts_bits).reshape(-1)
    clear = np.abs(om_ref / (4 * np.pi) - 0.5) > 1e-4
    assert np.array_equal(in_t[clear], in_ref[clear])


def test_cta_timeline_trace_and_launch_shapes(lb, prim, monkeypatch, tmp_path):
    """WN_TRACE_FILE (diagnostics): every CTA of the tiled kernels records {start, end, SM}; the dump lists, per batch, the two block
    levels, the tile plan (one CTA per tile) and the tile query, whose grid is runs of tiles plus single-tile CTAs at the light end
    (WN_TILE_RUN / WN_TILE_TAIL). Results do not depend on the run length, the tail or the two-lane dispatch."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("cta_timeline", os.path.join(os.path.dirname(__file__), "..", "tools", "cta_timeline.py"))
    tl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tl)
    monkeypatch.setenv("WN_TILE", "1")
    V, F = prim.generate_torus(5.0, 1.0, 64, 32)
    eng = lb.FastWindingNumber(V, F)
    lo, hi = prim.mesh_bbox(V)
    d = np.array((96, 80, 72), dtype=np.int64)  # 12 x 10 x 9 = 1080 tiles
    o = (lo - 0.3).astype(np.float32)
    s = ((hi - lo + 0.6) / d).astype(np.float32)
    ref = eng.query_grid(o, s, d, want_omega=True)[0]
    for run, tail, lanes in (("3", "0", "1"), ("4", "1", "1"), ("1", "2", "1"), ("2", "2", "2")):
        monkeypatch.setenv("WN_TILE_RUN", run)
        monkeypatch.setenv("WN_TILE_TAIL", tail)
        monkeypatch.setenv("WN_TILE_LANES", lanes)
        monkeypatch.setenv("WN_TILE_LANES_MIN", "1")
        path = tmp_path / f"trace_{run}_{tail}_{lanes}.bin"
        monkeypatch.setenv("WN_TRACE_FILE", str(path))
        om = eng.query_grid(o, s, d, want_omega=True)[0]
        monkeypatch.delenv("WN_TRACE_FILE")
        assert np.array_equal(om, ref), (run, tail, lanes)
        launches = tl.read_trace(str(path))
        tags = [t for t, _, _ in launches]
        assert tags.count(1) == tags.count(2) >= 1 and tags.count(11) == tags.count(12) == tags.count(1)
        assert sum(len(dd) for t, _, dd in launches if t == 1) == 1080  # one plan CTA per tile
        assert {ln for _, ln, _ in launches} == ({0, 1} if lanes == "2" else {0})
        for t, _, dd in launches:
            assert np.all(dd[:, 1] >= dd[:, 0]) and np.all(dd[:, 0] > 0), t  # every CTA wrote its start and end
        if lanes == "1":
            q = [dd for t, _, dd in launches if t == 2][0]
            r = int(run)
            slots = tl.summarise(launches)["launches"][-1]["sms"]  # >= 1 SM seen; the tail is sized from the device's resident CTAs
            assert slots >= 1
            if tail == "0" or r == 1:
                assert len(q) == -(-1080 // r)
            else:
                assert -(-1080 // r) < len(q) <= 1080  # some tiles went one per CTA
