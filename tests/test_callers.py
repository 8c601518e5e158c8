"""The reference's example callers as batched operations (lagrange_b200/callers.py).
CPU tier: the random stream against libstdc++'s own output, the per-facet probe points and the flip rule against a
literal per-facet restatement driven by the oracle. GPU tier: the same operations through the C-ABI engine."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_mt_golden():
    path = os.path.join(GOLDEN, "mt19937_uniform_float.txt")
    head = np.loadtxt(path, dtype=np.float32, max_rows=1)
    return head[:3], head[3:], np.loadtxt(path, dtype=np.float32, skiprows=1)


def flipped_copy(F, fraction, seed):
    rng = np.random.RandomState(seed)
    flip = rng.rand(len(F)) < fraction
    G = F.copy()
    G[flip, 0], G[flip, 1] = F[flip, 1], F[flip, 0]
    return G, flip


def per_facet_reference(V, F, solid_angle_one, epsilon=1e-2, threshold=0.8):
    """fix_orientation.cpp:86-108, one facet at a time, float32 vector arithmetic spelled out."""
    f32 = np.float32
    out = F.copy()
    crit = np.empty(len(F), dtype=np.float64)
    for ff in range(len(F)):
        aa, bb, cc = (V[F[ff, k]].astype(f32) for k in range(3))
        n = np.cross(bb - aa, cc - aa).astype(f32)
        n = (n / np.sqrt((n * n).sum(dtype=f32), dtype=f32)).astype(f32)
        bary = ((aa + bb + cc) / f32(3)).astype(f32)
        pp = (bary + f32(epsilon) * n).astype(f32)
        qq = (bary - f32(epsilon) * n).astype(f32)
        crit[ff] = float(f32(solid_angle_one(pp)) - f32(solid_angle_one(qq))) / (4.0 * 3.14159265358979323846)
        if crit[ff] > float(f32(threshold)):
            out[ff, 0], out[ff, 1] = F[ff, 1], F[ff, 0]
    return out, crit


def test_mt19937_stream_matches_libstdcxx():
    from lagrange_b200.callers import mt19937_uniform_float

    lo, hi, pts = load_mt_golden()
    raw = mt19937_uniform_float(3 * len(pts), 0.0, 1.0).reshape(-1, 3)
    mine = (raw * (hi - lo)[None, :] + lo[None, :]).astype(np.float32)
    assert np.array_equal(mine, pts)
    assert raw.min() >= 0.0 and raw.max() < 1.0


def test_probe_points_and_flip_rule_match_per_facet_restatement(prim, oracle_mod):
    from lagrange_b200 import callers

    V, F = prim.generate_torus(5, 1, 16, 10)
    G, flipped = flipped_copy(F, 0.2, 3)
    ref = oracle_mod.RefEngine(V, F)
    one = lambda p: ref.solid_angle(p.reshape(1, 3))[0]
    want_F, want_c = per_facet_reference(V, G, one)
    got_F, got_c, counts = callers.fix_orientation(V, G, ref)
    assert np.array_equal(got_F, want_F)
    assert np.array_equal(got_c, want_c)
    # the closed reference mesh is outward oriented: exactly the flipped facets are turned back
    assert np.array_equal(got_F, F)
    assert counts == {"positive": int(flipped.sum()), "negative": int((~flipped).sum()), "total": len(F)}


def test_degenerate_facet_is_left_alone(prim, oracle_mod):
    from lagrange_b200 import callers

    V, F = prim.generate_torus(5, 1, 12, 8)
    G = np.vstack([F, [[0, 0, 1]]]).astype(F.dtype)  # zero-area facet: NaN normal, criterion NaN, no flip
    got_F, crit, counts = callers.fix_orientation(V, G, oracle_mod.RefEngine(V, F))
    assert np.isnan(crit[-1]) and np.array_equal(got_F[-1], G[-1])
    assert counts["positive"] == 0 and counts["negative"] == len(F) and counts["total"] == len(F) + 1


def test_sample_points_in_mesh_with_oracle(prim, oracle_mod):
    from lagrange_b200 import callers

    V, F = prim.generate_torus(5, 1, 24, 12)
    lo, hi = V.min(axis=0), V.max(axis=0)
    ref = oracle_mod.RefEngine(V, F)
    pts = callers.sample_points_in_mesh(ref, lo, hi, 2000)
    assert pts.dtype == np.float32 and pts.shape[1] == 3
    # torus volume / bbox volume = 2 pi^2 R r^2 / ((2(R+r))^2 2r) = 0.343 for R=5, r=1 (facetted: a bit less)
    assert 0.27 < len(pts) / 2000 < 0.37
    rho = np.hypot(pts[:, 0], pts[:, 2]) - 5.0
    assert np.all(rho * rho + pts[:, 1] ** 2 < 1.0 + 1e-4)


@pytest.mark.gpu
def test_fix_orientation_gpu_equals_oracle(prim, oracle_mod):
    import lagrange_b200 as lb
    from lagrange_b200 import callers

    V, F = prim.generate_torus(5, 1, 60, 30)
    G, flipped = flipped_copy(F, 0.1, 7)
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    got_F, got_c, counts = callers.fix_orientation(V, G, eng)
    want_F, want_c, _ = callers.fix_orientation(V, G, oracle_mod.RefEngine(V, F))
    assert np.array_equal(got_F, F) and np.array_equal(got_F, want_F)
    assert counts["positive"] == int(flipped.sum())
    # different hierarchies (GPU LBVH vs the restatement's 4-ary SAH): the criterion agrees to both trees' truncation error
    assert np.abs(got_c - want_c).max() < 2e-2
    assert np.all(np.abs(np.abs(got_c) - 1.0) < 0.1)  # +-1 across a closed surface


@pytest.mark.gpu
def test_sample_points_in_mesh_gpu(prim, oracle_mod):
    import lagrange_b200 as lb
    from lagrange_b200 import callers

    V, F = prim.generate_torus(5, 1, 60, 30)
    lo, hi = V.min(axis=0), V.max(axis=0)
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    pts = callers.sample_points_in_mesh(eng, lo, hi, 10000)
    ref_pts = callers.sample_points_in_mesh(oracle_mod.RefEngine(V, F), lo, hi, 10000)
    # same stream, same predicate: the kept sets can differ only for samples within the trees' error of w = 0.5
    a = {tuple(p) for p in pts.tolist()}
    b = {tuple(p) for p in ref_pts.tolist()}
    assert len(a ^ b) <= 0.003 * 10000
    w = oracle_mod.exact64(V, F, pts[:500]) / (4 * np.pi)
    assert np.all(w > 0.45)


@pytest.mark.gpu
def test_mesh_distances_hausdorff_chamfer_against_the_brute_force(prim, oracle_mod):
    """compute_mesh_distances / compute_hausdorff / compute_chamfer (modules/bvh/src/compute_mesh_distances.cpp:45-166) over
    wn_closest_point, against the double-precision brute-force distance."""
    from lagrange_b200 import callers

    Va, Fa = prim.generate_torus(5.0, 1.0, 40, 20)
    Vb, Fb = prim.generate_torus(5.2, 0.8, 36, 18)
    Vb = (Vb + np.array([0.1, -0.05, 0.2], np.float32)).astype(np.float32)
    d = callers.compute_mesh_distances(Va, Vb, Fb)
    ref_ab = oracle_mod.distance64(Vb, Fb, Va)
    ref_ba = oracle_mod.distance64(Va, Fa, Vb)
    assert np.abs(d - ref_ab).max() < 1e-5
    assert callers.compute_hausdorff(Va, Fa, Vb, Fb) == pytest.approx(max(ref_ab.max(), ref_ba.max()), abs=1e-5)
    assert callers.compute_chamfer(Va, Fa, Vb, Fb) == pytest.approx((ref_ab ** 2).mean() + (ref_ba ** 2).mean(), rel=1e-5)
    assert np.array_equal(callers.compute_mesh_distances(Va, Vb, np.zeros((0, 3), np.int32)), np.zeros(len(Va), np.float32))
    # a mesh against itself: every vertex is on the target
    assert callers.compute_mesh_distances(Va, Va, Fa).max() < 1e-6
