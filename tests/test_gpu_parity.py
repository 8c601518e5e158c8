"""GPU tier (-m gpu): the CUDA path, called through the C-ABI, against the oracle on the same seeded inputs.

Bars (BASELINE.json north_star): is_inside bit-identical to the reference algorithm wherever |w_ref - 0.5| > 1e-3;
solid_angle within 1e-4 * 4 pi. With the oracle's topology imported (oracle-tree mode) the second bar is
met literally and the traversal counters are integer-identical; with the GPU-built LBVH the tree differs from the
reference's, so solid_angle is judged against exact64 with the restatement's own error as the yardstick.
"""
import glob
import os

import numpy as np
import pytest

from conftest import FOUR_PI, band_mask, small_config

pytestmark = pytest.mark.gpu

TOL_OMEGA = 1e-4 * FOUR_PI
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.fixture(scope="module")
def lb():
    import torch

    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    import lagrange_b200

    return lagrange_b200


# ---- K2 radix sort ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 31, 4096, 4097, 100003, 1 << 20])
def test_radix_sort_is_a_stable_sort(lb, n):
    import ctypes

    from lagrange_b200 import _capi

    rng = np.random.Generator(np.random.PCG64(n))
    L = _capi.lib()
    for dtype, fn, bits in ((np.uint64, L.wn_debug_sort_pairs_u64, 63), (np.uint32, L.wn_debug_sort_pairs_u32, 30)):
        # few distinct keys => many ties => stability matters
        keys = rng.integers(0, 1 << (bits if n < 5000 else 12), size=n, dtype=np.uint64).astype(dtype)
        if n > 10:
            keys[::7] = rng.integers(0, 1 << bits, size=len(keys[::7]), dtype=np.uint64).astype(dtype)
        vals = np.arange(n, dtype=np.uint32)
        k, v = keys.copy(), vals.copy()
        _capi.check(fn(ctypes.c_void_p(k.ctypes.data), ctypes.c_void_p(v.ctypes.data), n, 0, bits))
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(k, keys[order]) and np.array_equal(v, vals[order])


# ---- oracle-tree mode: same topology as the reference restatement --------------------------------------------------------
@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_oracle_tree_mode_matches_the_reference_algorithm(lb, oracle_mod, prim, cfg):
    V, F, q, lattice = small_config(prim, cfg)
    ref = oracle_mod.RefEngine(V, F)
    topo = ref.topology()
    eng = lb.FastWindingNumber(V, F, topology=topo, keep_build_data=True)
    info = eng.info
    assert info["width"] == 4 and info["num_entries"] == ref.num_nodes + len(F)
    # K4 moments: bit-identical to the oracle's stored lanes
    r23 = eng.debug_node_moments()
    bd = ref.boxdata()
    nI = ref.num_nodes
    sel = topo != -1
    nodes = np.where(topo >= 0, topo, nI - (topo + 2))
    assert np.array_equal(bd[sel], r23[nodes[sel]])
    assert np.array_equal(eng.debug_topology(), topo)
    for beta in (2.0, 3.0):
        o_ref, c_ref = ref.solid_angle(q, beta=beta, counters=True)
        o_gpu = eng.solid_angle(q, accuracy_scale=beta)
        assert np.abs(o_gpu - o_ref).max() < TOL_OMEGA, np.abs(o_gpu - o_ref).max() / FOUR_PI
        # same accepted set per point => the traversal counters agree with the oracle's. The device forms |r|^2 with
        # FMAs, so a point within an ulp of a node's threshold may flip: allow a 1e-5 relative slack, nothing more.
        st = eng.query_stats(q, accuracy_scale=beta)
        for got, want in zip((st["node_tests"], st["far_field_evals"], st["exact_triangles"]), (int(c) for c in c_ref)):
            assert abs(got - want) <= 3 + 1e-5 * want, (st, c_ref)
    ins_ref = ref.is_inside(q)
    ins_gpu = eng.is_inside(q)
    m = band_mask(ref.solid_angle(q) / FOUR_PI)
    assert np.array_equal(ins_gpu[m], ins_ref[m])
    assert np.mean(ins_gpu != ins_ref) < 1e-3
    if lattice is not None:
        om_g, ins_g = eng.query_grid(*lattice, want_omega=True, want_inside=True, tiling=False)
        assert np.array_equal(om_g, eng.solid_angle(q, tiling=False)) and np.array_equal(ins_g, eng.is_inside(q, tiling=False))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_golden_fixtures(lb, path):
    g = np.load(path)
    eng = lb.FastWindingNumber(g["V"], g["F"], topology=g["topology"])
    for beta in (2, 3):
        om = eng.solid_angle(g["Q"], accuracy_scale=float(beta))
        assert np.abs(om - g[f"ref_beta{beta}"]).max() < TOL_OMEGA
        st = eng.query_stats(g["Q"], accuracy_scale=float(beta))
        for got, want in zip((st["node_tests"], st["far_field_evals"], st["exact_triangles"]), g[f"cnt_beta{beta}"]):
            assert abs(got - int(want)) <= 3, (st, g[f"cnt_beta{beta}"])
    w = g["ref_beta2"] / FOUR_PI
    m = band_mask(w)
    assert np.array_equal(eng.is_inside(g["Q"])[m], g["inside_beta2"][m])
    ex = eng.exact_solid_angle(g["Q"])
    assert np.abs(ex - g["exact64"]).max() < TOL_OMEGA


# ---- GPU-built LBVH --------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_lbvh_build_matches_host_emulation_and_exact(lb, oracle_mod, emul_mod, prim, cfg):
    V, F, q, lattice = small_config(prim, cfg)
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh", keep_build_data=True)
    em = emul_mod.EmulEngine(V, F)
    # K1-K3: same Morton order and Karras topology as the sequential emulation of the same source
    assert np.array_equal(eng.debug_topology(), em.topology())
    # K4: moments bit-identical (unfused arithmetic on both sides)
    assert np.array_equal(eng.debug_node_moments(), em.ref23())
    assert eng.info["num_entries"] == em.num_entries and eng.info["max_depth"] == em.max_depth
    o_gpu = eng.solid_angle(q)
    o_em = em.solid_angle(q)
    assert np.abs(o_gpu - o_em).max() < TOL_OMEGA
    st = eng.query_stats(q)
    _, c_em = em.solid_angle(q, counters=True)
    assert abs(st["far_field_evals"] - int(c_em[1])) <= 1e-4 * int(c_em[1]) + 2  # FMA vs unfused test: rare boundary flips
    # accuracy: judged against exact64 with the reference restatement's own error as the yardstick
    ex = oracle_mod.exact64(V, F, q)
    ref = oracle_mod.RefEngine(V, F)
    o_ref = ref.solid_angle(q)
    err_gpu = np.abs(o_gpu - ex) / FOUR_PI
    err_ref = np.abs(o_ref - ex) / FOUR_PI
    assert err_gpu.max() < max(3.0 * err_ref.max(), 2e-3) and err_gpu.mean() < max(2.0 * err_ref.mean(), 5e-4)
    # is_inside: identical to the reference algorithm away from the 0.5 level set (band widened by both trees' error)
    ins_gpu, ins_ref = eng.is_inside(q), ref.is_inside(q)
    m = band_mask(o_ref / FOUR_PI, band=1e-3 + err_gpu.max() + err_ref.max())
    assert np.array_equal(ins_gpu[m], ins_ref[m])
    # beta sweep converges to exact (cfg5's sweep, reduced)
    for beta, bound in ((4.0, 1.5e-3), (8.0, 1e-4)):
        assert np.abs(eng.solid_angle(q, accuracy_scale=beta) - ex).max() / FOUR_PI < bound


@pytest.mark.parametrize("opts", [dict(leaf_size=4), dict(leaf_size=16), dict(morton_bits=30), dict(radius_mode="vertex"),
                                  dict(order=1), dict(order=0), dict(approximate_single_triangles=True)])
def test_lbvh_options_match_host_emulation(lb, emul_mod, prim, opts):
    V, F, q, _ = small_config(prim, 1)
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh", **opts)
    ekw = dict(opts)
    if "radius_mode" in ekw:
        ekw["radius_mode"] = 1
    if "approximate_single_triangles" in ekw:
        ekw["approx_single"] = 1 if ekw.pop("approximate_single_triangles") else 0
    em = emul_mod.EmulEngine(V, F, **ekw)
    assert eng.info["num_entries"] == em.num_entries
    assert np.abs(eng.solid_angle(q) - em.solid_angle(q)).max() < TOL_OMEGA


# ---- batching must not change per-point results ------------------------------------------------------------------------------
TOL_TILE = 3e-5 * FOUR_PI  # far-field interpolation of the tiled path (measured ~1e-5 * 4 pi), on top of which nothing else moves


def test_results_do_not_depend_on_batch_composition(lb, prim):
    """Generic traversal: a point's result is bit-identical whatever else is in its warp / batch."""
    V, F = prim.generate_torus(5, 1, 60, 30)
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    q = prim.uniform_points_in_bbox(*prim.mesh_bbox(V), 50000, seed=11)
    base = eng.solid_angle(q, tiling=False)  # Morton-sorted internally
    assert np.array_equal(eng.solid_angle(q, presorted=True, tiling=False), base)  # original (incoherent) order, no sort
    perm = np.random.Generator(np.random.PCG64(1)).permutation(len(q))
    assert np.array_equal(eng.solid_angle(q[perm], tiling=False)[np.argsort(perm)], base)
    for i in (0, 17, 49999):  # the reference's single-point signature
        assert eng.solid_angle(q[i]) == base[i]
        assert eng.is_inside(q[i]) == bool(base[i] >= np.float32(6.2831854820251465))
    assert np.array_equal(eng.solid_angle(q[:1000]), base[:1000])
    for qpl in ("1", "2"):
        os.environ["WN_QPL"] = qpl
        try:
            assert np.array_equal(eng.solid_angle(q, tiling=False), base)
        finally:
            del os.environ["WN_QPL"]
    # tiled path (sorted points grouped in tiles of 512): same accepted records, far field interpolated per tile
    tiled = eng.solid_angle(q)
    assert np.abs(tiled - base).max() < TOL_TILE
    assert np.array_equal(eng.is_inside(q), eng.is_inside(q, tiling=False)) or np.mean(eng.is_inside(q) != eng.is_inside(q, tiling=False)) < 1e-4
    # a caller that lies about coherence still gets correct answers (tiles overflow and fall back)
    assert np.abs(eng.solid_angle(q, presorted=True) - base).max() < TOL_TILE


def test_grid_overload_equals_points_and_slabs_tile(lb, prim):
    V, F = prim.generate_torus(5, 1, 50, 24)
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    o, s, d = prim.lattice_for_bbox(*prim.mesh_bbox(V), (37, 11, 29))
    P = prim.lattice_points(o, s, d)
    om, ins = eng.query_grid(o, s, d, want_omega=True, want_inside=True, tiling=False)
    assert np.array_equal(om, eng.solid_angle(P, tiling=False)) and np.array_equal(ins, eng.is_inside(P, tiling=False))
    parts = [eng.query_grid(o, s, d, z_range=(a, b), want_omega=True, tiling=False)[0] for a, b in ((0, 5), (5, 6), (6, 29))]
    assert np.array_equal(np.concatenate(parts), om)
    assert eng.query_grid(o, s, d, z_range=(4, 4))[1].size == 0
    ex_om, ex_ins = eng.exact_grid(o, s, d, want_omega=True, want_inside=True)
    assert np.array_equal(ex_om, eng.exact_solid_angle(P))
    st_g = eng.query_stats_grid(o, s, d)
    st_p = eng.query_stats(P)
    assert [st_g[k] for k in ("node_tests", "far_field_evals", "exact_triangles")] == [st_p[k] for k in ("node_tests", "far_field_evals", "exact_triangles")]
    # tiled lattice: whole lattice and z-slabs (slab cuts move the tile boundaries)
    os.environ["WN_TILE"] = "1"
    try:
        om_t, ins_t = eng.query_grid(o, s, d, want_omega=True, want_inside=True)
        assert np.abs(om_t - om).max() < TOL_TILE
        parts = [eng.query_grid(o, s, d, z_range=(a, b), want_omega=True)[0] for a, b in ((0, 5), (5, 6), (6, 29))]
        assert np.abs(np.concatenate(parts) - om).max() < TOL_TILE
        m = band_mask(om / FOUR_PI)
        assert np.array_equal(ins_t[m], ins[m])
        st_t = eng.query_stats_grid(o, s, d, tiling=True)
        assert st_t["far_field_evals"] < st_g["far_field_evals"] and st_t["node_tests"] < st_g["node_tests"]
        assert st_t["exact_triangles"] == st_g["exact_triangles"]  # near field is untouched by the tiling
    finally:
        del os.environ["WN_TILE"]


def test_strided_layers_tile_the_lattice(lb, prim):
    """Multi-GPU sharding primitive: layers r, r+N, ... of every rank together are the whole lattice, bit for bit."""
    V, F = prim.generate_subdivided_sphere("icosahedron", 4)
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    o, s, d = prim.lattice_for_bbox([-1.0] * 3, [1.0] * 3, (40, 24, 53))  # nz not a multiple of 8: partial last layer
    per = 40 * 24
    for tiling_env in ("0", "1"):
        os.environ["WN_TILE"] = tiling_env
        try:
            full_om, full_in = eng.query_grid(o, s, d, want_omega=True)
            for world in (2, 3, 8):
                seen = np.zeros(53, dtype=int)
                for rank in range(world):
                    planes = eng.strided_layer_planes(53, rank, world)
                    om, ins = eng.query_grid(o, s, d, want_omega=True, layers=(rank, world))
                    assert om.shape == (per * len(planes),)
                    for k, z in enumerate(planes):
                        assert np.array_equal(om[k * per:(k + 1) * per], full_om[z * per:(z + 1) * per])
                        assert np.array_equal(ins[k * per:(k + 1) * per], full_in[z * per:(z + 1) * per])
                        seen[z] += 1
                assert np.all(seen == 1)
        finally:
            del os.environ["WN_TILE"]


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("tree", ["lbvh", "oracle", "lbvh_leaf8"])
def test_tiled_path_matches_generic_traversal(lb, oracle_mod, prim, cfg, tree):
    """The tile plan may only change where the far field is summed, never which records a point accepts."""
    V, F, q, lattice = small_config(prim, cfg)
    if tree == "oracle":
        eng = lb.FastWindingNumber(V, F, topology=oracle_mod.RefEngine(V, F).topology())
    else:
        eng = lb.FastWindingNumber(V, F, hierarchy="lbvh", leaf_size=8 if tree == "lbvh_leaf8" else 1)
    os.environ["WN_TILE"] = "1"
    try:
        for beta in (2.0, 3.5):
            if lattice is not None:
                ref_om, ref_in = eng.query_grid(*lattice, want_omega=True, accuracy_scale=beta, tiling=False)
                om, ins = eng.query_grid(*lattice, want_omega=True, accuracy_scale=beta)
            else:
                ref_om, ref_in = eng.solid_angle(q, accuracy_scale=beta, tiling=False), eng.is_inside(q, accuracy_scale=beta, tiling=False)
                om, ins = eng.solid_angle(q, accuracy_scale=beta), eng.is_inside(q, accuracy_scale=beta)
            assert np.abs(om - ref_om).max() < TOL_TILE, np.abs(om - ref_om).max() / FOUR_PI
            m = band_mask(ref_om / FOUR_PI, band=1e-4)
            assert np.array_equal(ins[m], ref_in[m])
        for kappa in ("3", "16"):  # far-set distance criterion: accuracy must hold across the useful range
            os.environ["WN_KAPPA"] = kappa
            try:
                om = eng.query_grid(*lattice, want_omega=True)[0] if lattice is not None else eng.solid_angle(q)
                ref = eng.query_grid(*lattice, want_omega=True, tiling=False)[0] if lattice is not None else eng.solid_angle(q, tiling=False)
                assert np.abs(om - ref).max() < (1e-4 if kappa == "3" else 3e-5) * FOUR_PI
            finally:
                del os.environ["WN_KAPPA"]
    finally:
        del os.environ["WN_TILE"]


def test_tiled_path_survives_degenerate_tiles(lb, prim):
    V, F = prim.generate_torus(5, 1, 40, 20)
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    os.environ["WN_TILE"] = "1"
    try:
        # all points identical / collinear / containing NaN / a single point / exactly one tile + 1
        one = np.tile(np.array([[5.0, 0.1, 0.2]], np.float32), (1000, 1))
        assert np.abs(eng.solid_angle(one) - eng.solid_angle(one, tiling=False)).max() < TOL_TILE
        line = np.stack([np.linspace(-7, 7, 3000), np.zeros(3000), np.zeros(3000)], 1).astype(np.float32)
        assert np.abs(eng.solid_angle(line) - eng.solid_angle(line, tiling=False)).max() < TOL_TILE
        bad = prim.uniform_points_in_bbox(*prim.mesh_bbox(V), 513, seed=3)
        bad[100] = np.nan
        bad[200, 1] = np.inf
        a, b = eng.solid_angle(bad), eng.solid_angle(bad, tiling=False)
        ok = np.isfinite(b)
        assert ok.sum() >= 511 and np.abs(a[ok] - b[ok]).max() < TOL_TILE
        # a lattice much coarser than the mesh (tiles span the whole torus) and one much finer (tiles inside one triangle)
        for n, scale in ((16, 1.0), (64, 0.01)):
            lo, hi = prim.mesh_bbox(V)
            c = 0.5 * (lo + hi) + np.array([5.0, 0, 0]) * (scale < 1)
            o, s, d = prim.lattice_for_bbox(c - scale * (hi - lo), c + scale * (hi - lo), n)
            x = eng.query_grid(o, s, d, want_omega=True)[0]
            y = eng.query_grid(o, s, d, want_omega=True, tiling=False)[0]
            assert np.abs(x - y).max() < TOL_TILE
    finally:
        del os.environ["WN_TILE"]


# ---- K7 exact mode -----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 100, 5000])
def test_exact_mode_is_ground_truth(lb, oracle_mod, prim, n):
    V, F = prim.generate_torus(5, 1, 120, 50)  # 24 000 triangles: several chunks and tiles
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    q = prim.uniform_points_in_bbox(*prim.mesh_bbox(V), n, seed=n)
    ex = oracle_mod.exact64(V, F, q)
    om = eng.exact_solid_angle(q)
    assert np.abs(om - ex).max() < 2e-5 * FOUR_PI
    m = band_mask(ex / FOUR_PI)
    assert np.array_equal(eng.exact_is_inside(q).astype(bool)[m], (ex / FOUR_PI > 0.5)[m])


# ---- edge cases the reference's surface defines ------------------------------------------------------------------------------
def test_edge_cases(lb, oracle_mod):
    empty = lb.FastWindingNumber(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32), hierarchy="lbvh")
    q = np.array([[0, 0, 0], [1, 2, 3]], dtype=np.float32)
    assert np.all(empty.solid_angle(q) == 0) and not empty.is_inside(q).any() and np.all(empty.exact_solid_angle(q) == 0)
    assert empty.solid_angle(np.zeros((0, 3), np.float32)).shape == (0,)
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [2, 2, 2]], dtype=np.float32)
    F = np.array([[0, 1, 2], [0, 1, 2], [0, 1, 1], [4, 4, 4], [0, 2, 3], [0, 3, 1], [1, 3, 2], [0, 2, 1]], dtype=np.int32)
    qq = np.array([[0.1, 0.1, 0.1], [3, 3, 3], [0, 0, 0], [1, 0, 0], [0.25, 0.25, 0], [-1, 0.5, 0.2]], dtype=np.float32)
    ex = oracle_mod.exact64(V, F, qq)
    for kw in ({}, {"leaf_size": 4}, {"topology": oracle_mod.RefEngine(V, F).topology()}):
        eng = lb.FastWindingNumber(V, F, hierarchy="lbvh", **kw)
        assert np.abs(eng.solid_angle(qq, accuracy_scale=50.0) - ex).max() < 1e-4
        assert np.abs(eng.exact_solid_angle(qq) - ex).max() < 1e-5
    one = lb.FastWindingNumber(V, F[4:5], hierarchy="lbvh")
    assert abs(one.solid_angle([0.1, 0.1, 0.1]) - oracle_mod.exact64(V, F[4:5], qq[:1])[0]) < 1e-5
    with pytest.raises(lb.Error, match="vertex index"):
        lb.FastWindingNumber(V, np.array([[0, 1, 7]], dtype=np.int32), hierarchy="lbvh")
    with pytest.raises(lb.Error, match="topology"):
        lb.FastWindingNumber(V, F[:2], topology=np.array([[-2, -2, -1, -1]], dtype=np.int32))
    with pytest.raises(lb.Error):
        lb.FastWindingNumber(V, F, hierarchy="lbvh").solid_angle(np.zeros((4, 2), np.float32))
    # non-finite queries must not hang or poison their neighbours
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    bad = np.array([[np.nan, 0, 0], [0.1, 0.1, 0.1], [np.inf, 0, 0]], dtype=np.float32)
    r = eng.solid_angle(bad)
    assert abs(r[1] - eng.solid_angle([0.1, 0.1, 0.1])) == 0


def test_device_pointers_and_pack_roundtrip(lb, prim):
    import torch

    V, F = prim.generate_subdivided_sphere("icosahedron", 4)
    eng = lb.FastWindingNumber(torch.from_numpy(V).cuda(), torch.from_numpy(F).cuda(), hierarchy="lbvh")  # device-resident mesh
    q = prim.uniform_points_in_bbox([-1.2] * 3, [1.2] * 3, 20000, seed=5)
    host = eng.solid_angle(q)
    dq = torch.from_numpy(q).cuda()
    dev = eng.solid_angle(dq)
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), host)
    assert np.array_equal(eng.is_inside(dq).cpu().numpy(), eng.is_inside(q))
    o, s, d = prim.lattice_for_bbox([-1.0] * 3, [1.0] * 3, 24)
    om_d, ins_d = eng.query_grid(o, s, d, want_omega=True, device_output=True)
    om_h, ins_h = eng.query_grid(o, s, d, want_omega=True)
    assert om_d.is_cuda and np.array_equal(om_d.cpu().numpy(), om_h) and np.array_equal(ins_d.cpu().numpy(), ins_h)
    # inside count ~ sphere volume / cell volume
    cell = float(np.prod(s))
    assert abs(ins_h.sum() * cell - 4.0 / 3.0 * np.pi) < 0.15
    # packed tree: host and device round trips give engines with identical answers
    blob = eng.pack()
    clone = lb.FastWindingNumber.from_packed(blob)
    assert np.array_equal(clone.solid_angle(q), host)
    dblob = torch.empty(eng.packed_size(), dtype=torch.uint8, device="cuda")
    eng.pack(out=dblob)
    clone2 = lb.FastWindingNumber.from_packed(dblob)
    assert np.array_equal(clone2.solid_angle(q), host)
    with pytest.raises(lb.Error, match="magic"):
        lb.FastWindingNumber.from_packed(np.zeros(4096, np.uint8))


def test_cpp_host_layer():
    """The drop-in C++ class (include/lagrange/winding/FastWindingNumber.h) through the reference's benchmark protocol."""
    import subprocess

    from lagrange_b200 import build

    build.build_all()
    r = subprocess.run([build.CPP_TEST], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout


# ---- full-size properties (BASELINE configs at their real sizes) ----------------------------------------------------------------
def test_cfg1_full_size_against_the_oracle(lb, oracle_mod, prim):
    V, F = prim.config_mesh(1)
    _, lattice = prim.config_queries(1, V, F)
    ref = oracle_mod.RefEngine(V, F)
    eng_ref_tree = lb.FastWindingNumber(V, F, topology=ref.topology())
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    ins_ref, om_ref = ref.grid(*lattice, want_omega=True)
    om_t, ins_t = eng_ref_tree.query_grid(*lattice, want_omega=True)
    assert np.abs(om_t - om_ref).max() < TOL_OMEGA
    m = band_mask(om_ref / FOUR_PI)
    assert np.array_equal(ins_t[m], ins_ref[m])
    om_l, ins_l = eng.query_grid(*lattice, want_omega=True)
    mismatch = int((ins_l[m] != ins_ref[m]).sum())
    assert mismatch <= 5, mismatch  # different tree: agreement is statistical near the surface
    sub = slice(None, None, 97)
    P = prim.lattice_points(*lattice)[sub]
    ex = oracle_mod.exact64(V, F, P)
    assert np.abs(om_l[sub] - ex).max() / FOUR_PI < 1.5e-2
    assert np.abs(eng.exact_solid_angle(P) - ex).max() / FOUR_PI < 2e-5


def test_cfg2_full_size_sphere_properties(lb, prim):
    import torch

    V, F = prim.config_mesh(2)
    assert len(F) == 1310720
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    info = eng.info
    assert info["num_entries"] == 2 * len(F) - 1 and info["build_ms"] > 0
    kind, (o, s, d) = prim.config_queries(2, V, F)
    # one 64-layer slab of the 512^3 lattice through the middle; results stay on the device
    z0, z1 = 224, 288
    ins = eng.query_grid(o, s, d, z_range=(z0, z1), device_output=True)[1].view(z1 - z0, 512, 512)
    # analytic: the unit sphere. Points farther than one cell diagonal from the surface must be classified exactly.
    ax = torch.tensor(o[0], device="cuda") + torch.tensor(s[0], device="cuda") * (torch.arange(512, device="cuda", dtype=torch.float32) + 0.5)
    zz, yy, xx = torch.meshgrid(ax[z0:z1], ax, ax, indexing="ij")
    rad = torch.sqrt(xx * xx + yy * yy + zz * zz)
    clear = (rad - 1.0).abs() > 0.01
    assert torch.equal(ins.bool()[clear], (rad < 1.0)[clear])
    # symmetry of the lattice and of the icosphere under the central inversion z -> -z of this slab
    assert (ins != ins.flip(0)).float().mean() < 1e-4


# ---- K3': balanced k-d hierarchy (wn_options.hierarchy = WN_HIERARCHY_KD) ----------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("hierarchy", ["kd", "kd_sah"])
@pytest.mark.parametrize("cfg,leaf", [(1, 1), (3, 1), (1, 4)])
def test_kd_hierarchy_equals_host_emulation(prim, emul_mod, cfg, leaf, hierarchy):
    import lagrange_b200 as lb

    V, F, q, _ = small_config(prim, cfg)
    eng = lb.FastWindingNumber(V, F, hierarchy=hierarchy, keep_build_data=True, leaf_size=leaf)
    em = emul_mod.EmulEngine(V, F, hierarchy=hierarchy, leaf_size=leaf)
    assert np.array_equal(eng.debug_topology(), em.topology())
    # same topology, same moments (unfused build arithmetic), same folded records: the traversal differs only by the device's
    # rsqrt / FMA contraction in the exact-triangle term
    got = eng.solid_angle(q[:20000], tiling=False)
    want = em.solid_angle(q[:20000])
    assert np.abs(got - want).max() < 2e-5 * FOUR_PI


@pytest.mark.gpu
def test_kd_hierarchy_error_class_and_tiled_path(prim, oracle_mod):
    import lagrange_b200 as lb

    V, F, q, lattice = small_config(prim, 1)
    kd = lb.FastWindingNumber(V, F, hierarchy="kd")
    lbvh = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    w_exact = oracle_mod.exact64(V, F, q) / FOUR_PI
    e_kd = np.abs(kd.solid_angle(q) / FOUR_PI - w_exact)
    e_lb = np.abs(lbvh.solid_angle(q) / FOUR_PI - w_exact)
    assert e_kd.max() < 1.5 * e_lb.max() and e_kd.mean() < 1.5 * e_lb.mean()
    clear = band_mask(w_exact, 2.0 * e_kd.max())
    assert np.array_equal(kd.is_inside(q)[clear].astype(bool), (w_exact > 0.5)[clear])
    # tiled vs generic on the k-d tree: same accepted records, far set interpolated
    o, s, d = lattice
    om_t = kd.solid_angle_grid(o, s, d, tiling=True)
    om_g = kd.solid_angle_grid(o, s, d, tiling=False)
    assert np.abs(om_t - om_g).max() < 3e-5 * FOUR_PI
    # leaf_size > 1 collapses subtrees of the balanced tree like it does for the LBVH
    kd8 = lb.FastWindingNumber(V, F, hierarchy="kd", leaf_size=8)
    e8 = np.abs(kd8.solid_angle(q) / FOUR_PI - w_exact)
    assert e8.max() < 1.5 * e_lb.max()


@pytest.mark.gpu
def test_kd_hierarchy_degenerate_inputs(prim):
    import lagrange_b200 as lb

    # all centroids identical (stacked copies of one triangle), two triangles, three triangles
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32)
    for copies in (2, 3, 37):
        F = np.tile(np.array([[0, 1, 2]], dtype=np.int32), (copies, 1))
        ref = lb.FastWindingNumber(V, F, hierarchy="lbvh")
        q = np.array([[0.2, 0.2, 0.5], [0.2, 0.2, -0.5], [3, 3, 3]], dtype=np.float32)
        for hierarchy in ("kd", "kd_sah"):
            eng = lb.FastWindingNumber(V, F, hierarchy=hierarchy)
            assert np.allclose(eng.solid_angle(q), ref.solid_angle(q), atol=1e-5 * copies)


@pytest.mark.gpu
def test_calls_on_different_streams_share_the_engine_safely(prim):
    """Per-engine scratch (tile plans, staging) is reused by every call; consecutive calls on different CUDA streams are
    ordered by the library (an event between them), so device-resident results do not depend on the stream pattern."""
    import torch

    import lagrange_b200 as lb

    V, F = prim.generate_torus(5, 1, 120, 60)
    eng = lb.FastWindingNumber(V, F, hierarchy="kd", leaf_size=4)
    lattices = [prim.lattice_for_bbox(*prim.mesh_bbox(V), (96 + 8 * k, 40, 96), inflate=0.05 + 0.01 * k) for k in range(4)]
    want = [eng.query_grid(o, s, d, want_omega=True, device_output=True) for o, s, d in lattices]
    want = [(om.clone(), ins.clone()) for om, ins in want]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    got = []
    for rep in range(3):
        got.clear()
        for k, (o, s, d) in enumerate(lattices):
            with torch.cuda.stream(streams[k & 1]):
                got.append(eng.query_grid(o, s, d, want_omega=True, device_output=True))
        torch.cuda.synchronize()
        for (om, ins), (wom, wins) in zip(got, want):
            assert torch.equal(om, wom) and torch.equal(ins, wins)
