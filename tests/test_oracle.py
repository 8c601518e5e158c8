"""CPU tier: the oracle against implementation-independent known answers and its own frozen fixtures.

The reference pins nothing for this path ('parity unpinned'), so the oracle is anchored on
analytic facts of the solid angle (section 8(c) 'known answers the new repo can pin itself').
"""
import glob
import os

import numpy as np
import pytest

from conftest import FOUR_PI, band_mask, small_config

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def test_octant_triangle_is_one_eighth_of_the_sphere(oracle_mod):
    V = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float32)
    F = np.array([[0, 1, 2]], dtype=np.int32)
    om = oracle_mod.exact64(V, F, np.zeros((1, 3), dtype=np.float32))[0]
    assert abs(om - FOUR_PI / 8) < 1e-12  # normal points away from the origin => positive
    assert abs(oracle_mod.exact64(V, F[:, ::-1].copy(), np.zeros((1, 3), dtype=np.float32))[0] + FOUR_PI / 8) < 1e-12
    ref = oracle_mod.RefEngine(V, F)
    assert abs(ref.solid_angle(np.zeros((1, 3)))[0] - FOUR_PI / 8) < 1e-5


def test_closed_mesh_is_one_inside_zero_outside(oracle_mod, prim):
    V, F = prim.generate_subdivided_sphere("icosahedron", 3)
    q = np.array([[0, 0, 0], [0.3, -0.2, 0.5], [2, 0, 0], [0, -5, 1]], dtype=np.float32)
    w = oracle_mod.exact64(V, F, q) / FOUR_PI
    assert np.allclose(w, [1, 1, 0, 0], atol=1e-12)
    # flipped orientation => -1 ; two nested copies => 2
    assert np.allclose(oracle_mod.exact64(V, F[:, ::-1].copy(), q[:1]) / FOUR_PI, -1, atol=1e-12)
    V2 = np.concatenate([V, 0.5 * V]).astype(np.float32)
    F2 = np.concatenate([F, F + len(V)]).astype(np.int32)
    assert np.allclose(oracle_mod.exact64(V2, F2, q[:1]) / FOUR_PI, 2, atol=1e-12)
    ref = oracle_mod.RefEngine(V2, F2)
    assert np.allclose(ref.solid_angle(q[:1]) / FOUR_PI, 2, atol=5e-3)


def test_hemisphere_from_its_centre_is_one_half(oracle_mod, prim):
    V, F = prim.generate_subdivided_sphere("octahedron", 4)
    c = V[F].mean(axis=1)
    Fh = F[c[:, 2] > 0]  # the octahedron's equator is an edge loop, so this is exactly the upper hemisphere
    w = oracle_mod.exact64(V, Fh, np.zeros((1, 3), dtype=np.float32))[0] / FOUR_PI
    assert abs(w - 0.5) < 1e-12


def test_query_on_a_vertex_or_in_plane_contributes_zero(oracle_mod):
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32)
    F = np.array([[0, 1, 2]], dtype=np.int32)
    q = np.array([[0, 0, 0], [1, 0, 0], [0.25, 0.25, 0], [3, 3, 0]], dtype=np.float32)  # vertex, vertex, coplanar x2
    assert np.all(oracle_mod.exact64(V, F, q) == 0)
    assert np.all(oracle_mod.exact32(V, F, q) == 0)
    assert np.all(oracle_mod.RefEngine(V, F).solid_angle(q) == 0)


def test_inside_predicate_is_the_float_threshold_just_above_two_pi(oracle_mod):
    """FastWindingNumber.cpp:66 in double == omega >= 6.2831854820251465f."""
    thr = np.float32(6.2831854820251465)
    below = np.nextafter(thr, np.float32(0))
    assert oracle_mod.inside_predicate(float(thr)) and not oracle_mod.inside_predicate(float(below))
    x = thr
    for _ in range(50):  # a neighbourhood of floats on both sides
        assert oracle_mod.inside_predicate(float(x))
        x = np.nextafter(x, np.float32(100))
    x = below
    for _ in range(50):
        assert not oracle_mod.inside_predicate(float(x))
        x = np.nextafter(x, np.float32(0))
    for v in (0.0, -7.0, 6.0, 6.5, 12.6, np.inf, -np.inf):
        assert oracle_mod.inside_predicate(v) == (np.float32(v) >= thr)
    assert not oracle_mod.inside_predicate(float("nan"))


def test_reference_tree_is_a_partition_of_the_triangles(oracle_mod, prim):
    V, F = prim.generate_torus(5, 1, 40, 20)
    ref = oracle_mod.RefEngine(V, F)
    topo = ref.topology()
    tris = -(topo[topo <= -2] + 2)
    assert sorted(tris.tolist()) == list(range(len(F)))
    internal = topo[topo >= 0]
    assert sorted(internal.tolist()) == list(range(1, ref.num_nodes))  # every non-root node referenced exactly once
    # empties are trailing and only small nodes have them
    for row in topo:
        seen_empty = False
        for c in row:
            if c == -1:
                seen_empty = True
            else:
                assert not seen_empty


@pytest.mark.parametrize("cfg", [1, 2, 3, 5])
def test_restatement_error_and_beta_convergence(oracle_mod, prim, cfg):
    V, F, q, _ = small_config(prim, cfg)
    q = q[:: max(1, len(q) // 4000)]
    ex = oracle_mod.exact64(V, F, q) / FOUR_PI
    ref = oracle_mod.RefEngine(V, F)
    errs = []
    for beta in (2.0, 4.0, 8.0):
        w = ref.solid_angle(q, beta=beta).astype(np.float64) / FOUR_PI
        errs.append(np.abs(w - ex).max())
    assert errs[0] < 3e-2 and errs[1] < errs[0] and errs[2] < 2e-4, errs  # order-2 expansion: error falls fast with beta
    inside = ref.is_inside(q)
    m = band_mask(ex, band=errs[0] + 1e-3)
    assert np.array_equal(inside[m].astype(bool), ex[m] > 0.5)


def test_counters_define_the_work(oracle_mod, prim):
    V, F = prim.generate_torus(5, 1, 30, 16)
    ref = oracle_mod.RefEngine(V, F)
    q = prim.uniform_points_in_bbox(*prim.mesh_bbox(V), 500, seed=7)
    _, cnt2 = ref.solid_angle(q, beta=2.0, counters=True)
    _, cnt4 = ref.solid_angle(q, beta=4.0, counters=True)
    T, A, E = [int(x) for x in cnt2]
    assert T > A > 0 and E > 0 and all(int(a) > int(b) for a, b in zip(cnt4, cnt2))
    # beta -> huge: everything descends, every triangle is evaluated exactly for every query
    om, cnt = ref.solid_angle(q[:20], beta=1e6, counters=True)
    assert int(cnt[1]) == 0 and int(cnt[2]) == 20 * len(F)
    assert np.abs(om - oracle_mod.exact32(V, F, q[:20])).max() < 2e-4


def test_empty_mesh_gives_zero(oracle_mod):
    ref = oracle_mod.RefEngine(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32))
    assert ref.num_nodes == 0
    assert np.all(ref.solid_angle(np.ones((3, 3), np.float32)) == 0)
    assert not ref.is_inside(np.ones((3, 3), np.float32)).any()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_golden_fixtures_pin_the_oracle(oracle_mod, path):
    g = np.load(path)
    V, F, Q = g["V"], g["F"], g["Q"]
    assert np.array_equal(oracle_mod.exact64(V, F, Q), g["exact64"])
    ref = oracle_mod.RefEngine(V, F)
    assert np.array_equal(ref.topology(), g["topology"])
    for beta in (2, 3):
        om, cnt = ref.solid_angle(Q, beta=float(beta), counters=True)
        assert np.array_equal(om, g[f"ref_beta{beta}"])
        assert np.array_equal(cnt, g[f"cnt_beta{beta}"])
    assert np.array_equal(ref.is_inside(Q), g["inside_beta2"])


def test_grid_entry_point_matches_points(oracle_mod, prim):
    V, F = prim.generate_torus(5, 1, 24, 12)
    o, s, d = prim.lattice_for_bbox(*prim.mesh_bbox(V), 12)
    ref = oracle_mod.RefEngine(V, F)
    P = prim.lattice_points(o, s, d)
    ins, om = ref.grid(o, s, d, want_omega=True)
    assert np.array_equal(om, ref.solid_angle(P)) and np.array_equal(ins, ref.is_inside(P))
    ins2 = ref.grid(o, s, d, first=5, stride=7)
    assert np.array_equal(ins2, ins[5::7])


def test_simd_lane_evaluation_is_bit_identical_to_the_scalar_loop(oracle_mod, prim):
    """The 4 child lanes of a node are evaluated 4-wide (SSE) like upstream's v4uf path; same operations, same order."""
    V, F = prim.config_mesh(3, small=True)
    o, s, d = prim.lattice_for_bbox(*prim.mesh_bbox(V), 24)
    P = prim.lattice_points(o, s, d)
    try:
        for order in (0, 1, 2):
            ref = oracle_mod.RefEngine(V, F, order=order)
            oracle_mod.set_simd_lanes(True)
            a, ca = ref.solid_angle(P, counters=True)
            oracle_mod.set_simd_lanes(False)
            b, cb = ref.solid_angle(P, counters=True)
            assert np.array_equal(a, b) and np.array_equal(ca, cb)
    finally:
        oracle_mod.set_simd_lanes(True)
