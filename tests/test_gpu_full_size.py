"""GPU tier: the BASELINE.json configs at their FULL sizes, checked through the oracle where it
finishes in seconds on the box's host cores and through size-independent properties elsewhere."""
import numpy as np
import pytest

from conftest import FOUR_PI, band_mask

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1200)]

TOL_OMEGA = 1e-4 * FOUR_PI


@pytest.fixture(scope="module")
def lb():
    import torch

    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    import lagrange_b200

    return lagrange_b200


def test_cfg2_full_size_oracle_tree_parity(lb, oracle_mod, prim):
    """1 310 720 triangles, 512^3 lattice (134 M queries): the tiled CUDA path on the reference restatement's tree against the
    restatement itself, every lattice point."""
    V, F = prim.config_mesh(2)
    _, (o, s, d) = prim.config_queries(2, V, F)
    ref = oracle_mod.RefEngine(V, F)
    eng = lb.FastWindingNumber(V, F, topology=ref.topology())
    om, ins = eng.query_grid(o, s, d, want_omega=True, want_inside=True)
    ins_ref, om_ref = ref.grid(o, s, d, want_omega=True)
    diff = np.abs(om - om_ref)
    assert diff.max() < TOL_OMEGA, diff.max() / FOUR_PI
    m = band_mask(om_ref / FOUR_PI)
    assert np.array_equal(ins[m], ins_ref[m])
    assert int((ins != ins_ref).sum()) <= 8  # only inside the 1e-3 band around w = 1/2
    # the product build (LBVH) on the same lattice: classification agrees away from the surface shell
    eng2 = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    ins2 = eng2.query_grid(o, s, d)[1]
    assert np.array_equal(ins2[m], ins_ref[m]) or int((ins2[m] != ins_ref[m]).sum()) <= 16
    # analytic: the unit sphere
    assert abs(float(ins2.sum()) * float(np.prod(s)) - 4.0 / 3.0 * np.pi) < 2e-3


def test_cfg3_full_size_open_soup(lb, oracle_mod, prim):
    """200 k-triangle torus with holes, duplicates and flips, 256^3 lattice: generalized-WN robustness."""
    V, F = prim.config_mesh(3)
    _, (o, s, d) = prim.config_queries(3, V, F)
    assert tuple(d) == (256, 256, 256) and 150_000 < len(F) < 210_000
    ref = oracle_mod.RefEngine(V, F)
    ins_ref, om_ref = ref.grid(o, s, d, want_omega=True)
    w_ref = om_ref / FOUR_PI
    assert 1e-5 < np.mean(np.abs(w_ref - 0.5) < 1e-2) < 0.5  # an open soup really has points near the 1/2 level set
    # oracle-tree mode: the bar of BASELINE.json, literally
    eng_t = lb.FastWindingNumber(V, F, topology=ref.topology())
    om_t, ins_t = eng_t.query_grid(o, s, d, want_omega=True)
    assert np.abs(om_t - om_ref).max() < TOL_OMEGA
    m = band_mask(w_ref)
    assert np.array_equal(ins_t[m], ins_ref[m])
    # product build: different tree => both are ~1e-3 from the exact winding number; agreement outside the widened band
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    om, ins = eng.query_grid(o, s, d, want_omega=True)
    sub = slice(None, None, 4099)
    P = prim.lattice_points(o, s, d)[sub]
    ex = eng.exact_solid_angle(P)
    err_gpu = np.abs(om[sub] - ex).max() / FOUR_PI
    err_ref = np.abs(om_ref[sub] - ex).max() / FOUR_PI
    assert err_gpu < max(2.5 * err_ref, 5e-3), (err_gpu, err_ref)
    wide = band_mask(w_ref, band=1e-3 + 2.0 * (err_gpu + err_ref))
    assert np.array_equal(ins[wide], ins_ref[wide])
    assert np.mean(ins != ins_ref) < 2e-3


def test_cfg4_full_size_deep_traversal(lb, oracle_mod, prim):
    """8 388 608 triangles, 64 M near-surface jittered queries in random order (Morton sort + tiled point path)."""
    V, F = prim.config_mesh(4)
    assert len(F) == 8 * 4**10
    q = prim.near_surface_points(V, F, 64 << 20, seed=0xC0FFEE04)
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    info = eng.info
    assert info["num_entries"] == 2 * len(F) - 1
    ins = eng.is_inside(q)
    rad = np.linalg.norm(q.astype(np.float64), axis=1)
    # the octasphere's facets lie within ~1e-6 of the unit sphere; points farther than 1e-4 from it are unambiguous
    clear = np.abs(rad - 1.0) > 1e-4
    assert clear.mean() > 0.9
    assert np.array_equal(ins[clear].astype(bool), rad[clear] < 1.0)
    # a 1 M subsample: same answers when evaluated alone, unsorted, untiled (batch composition is irrelevant) ...
    sub = np.random.Generator(np.random.PCG64(4)).choice(len(q), 1 << 20, replace=False)
    om_sub = eng.solid_angle(q[sub], presorted=True, tiling=False)
    om_all = eng.solid_angle(q[sub])
    assert np.abs(om_sub - om_all).max() < 3e-5 * FOUR_PI
    # ... and against the reference restatement's own tree on 200 k of them
    small = sub[:200_000]
    ref = oracle_mod.RefEngine(V, F)
    w_ref = ref.solid_angle(q[small]) / FOUR_PI
    far = band_mask(w_ref, band=2e-2)
    assert np.array_equal(ins[small][far], ref.is_inside(q[small])[far])


def test_cfg5_full_size_exact_vs_tree_sweep(lb, oracle_mod, prim):
    """100 k triangles x 2^24 uniform points: exact brute-force mode vs the tree at beta in {1, 1.5, 2, 3, 4, 6, 8}."""
    import torch

    V, F = prim.config_mesh(5)
    assert len(F) == 100_000
    q = prim.uniform_points_in_bbox(*prim.mesh_bbox(V), 1 << 24, seed=0xC0FFEE05)
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    dq = torch.from_numpy(q).cuda()
    exact = eng.exact_solid_angle(dq)  # 1.7e12 triangle-point pairs on the device
    ex64 = oracle_mod.exact64(V, F, q[:: 1 << 14])
    assert np.abs(exact[:: 1 << 14].cpu().numpy() - ex64).max() < 3e-5 * FOUR_PI
    errs = {}
    for beta in (1.0, 1.5, 2.0, 3.0, 4.0, 6.0, 8.0):
        om = eng.solid_angle(dq, accuracy_scale=beta)
        errs[beta] = float((om - exact).abs().max().item()) / FOUR_PI
    print("cfg5 max |w_tree - w_exact| per beta:", errs)
    assert errs[2.0] < 2e-2 and errs[8.0] < 1e-4
    assert all(errs[b] >= errs[c] * 0.8 for b, c in zip((1.0, 1.5, 2.0, 3.0, 4.0, 6.0), (1.5, 2.0, 3.0, 4.0, 6.0, 8.0)))
    ins = eng.is_inside(dq)
    ins_exact = eng.exact_is_inside(dq)
    w = exact / FOUR_PI
    m = (w - 0.5).abs() > 1e-3 + errs[2.0]
    assert torch.equal(ins[m], ins_exact[m])


@pytest.mark.parametrize("hierarchy,leaf", [("reference", 1), ("kd_sah", 4)])
def test_cfg2_full_size_bench_configuration(lb, oracle_mod, prim, hierarchy, leaf):
    """What bench.py measures, at full size (512^3 lattice, tiled path). bench.py's default is the reference hierarchy; kd_sah with
    4-triangle leaves (round 1's bench configuration) is kept as the alternative. Against the restatement on EVERY lattice point
    with the strict 1e-3 band of BASELINE.json (the mismatch count is printed and asserted), against the exact winding number on
    a sample, and against the analytic volume of the unit sphere. Parity is against this repo's restatement, not the upstream binary."""
    V, F = prim.config_mesh(2)
    _, (o, s, d) = prim.config_queries(2, V, F)
    eng = lb.FastWindingNumber(V, F, hierarchy=hierarchy, leaf_size=leaf)
    om, ins = eng.query_grid(o, s, d, want_omega=True, want_inside=True)
    # same engine, per-point traversal instead of tiles: the far-set interpolation is the only difference
    om_g = eng.query_grid(o, s, d, want_omega=True, want_inside=False, tiling=False)[0]
    assert np.abs(om - om_g).max() < 3e-5 * FOUR_PI
    # strided sharding (what a rank of an 8-GPU run evaluates) returns exactly the corresponding planes
    planes = eng.strided_layer_planes(int(d[2]), 3, 8)
    part = eng.query_grid(o, s, d, layers=(3, 8))[1].reshape(len(planes), int(d[1]), int(d[0]))
    assert np.array_equal(part, ins.reshape(int(d[2]), int(d[1]), int(d[0]))[planes])
    # exact winding number on a sample of the lattice (brute force on the GPU, itself checked against exact64 elsewhere)
    sub = slice(None, None, 65537)
    P = prim.lattice_points(o, s, d)[sub]
    w_ex = eng.exact_solid_angle(P) / FOUR_PI
    err = np.abs(om[sub] / FOUR_PI - w_ex)
    assert err.max() < 6e-3 and err.mean() < 1e-3, (err.max(), err.mean())
    # the restatement on every lattice point, strict band
    ref = oracle_mod.RefEngine(V, F)
    ins_ref, om_ref = ref.grid(o, s, d, want_omega=True)
    strict = band_mask(om_ref / FOUR_PI, 1e-3)
    mism = int((ins[strict] != ins_ref[strict]).sum())
    domega = float(np.abs(om - om_ref).max()) / FOUR_PI
    print(f"cfg2 {hierarchy}/leaf {leaf}: is_inside mismatches vs the restatement outside |w-0.5|<=1e-3: {mism} of {int(strict.sum())} "
          f"(anywhere: {int((ins != ins_ref).sum())}); max |dOmega| = {domega:.2e} * 4pi; error vs exact: max {err.max():.2e} mean {err.mean():.2e}")
    if hierarchy == "reference":
        assert mism == 0 and domega < 1e-4
    else:
        # a different tree: both trees are ~1e-3 from the exact winding number, so the strict band is not a guarantee; the count
        # is reported and bounded here
        assert mism <= 64
        err_ref = np.abs(om_ref[sub] / FOUR_PI - w_ex).max()
        wide = band_mask(om_ref / FOUR_PI, band=1e-3 + 2.0 * (err.max() + err_ref))
        assert np.array_equal(ins[wide], ins_ref[wide])
    assert abs(float(ins.sum()) * float(np.prod(s)) - 4.0 / 3.0 * np.pi) < 2e-3


@pytest.mark.parametrize("cfg", [3, 4, 5])
def test_kd_sah_leaf4_full_size_strict_band_counts(lb, oracle_mod, prim, cfg):
    """the alternative hierarchy (kd_sah, 4-triangle leaves) at full size on cfg3 / cfg4 / cfg5 — strict-band
    mismatch counts against the restatement (bounded samples for the point-set configs), printed and bounded."""
    V, F = prim.config_mesh(cfg)
    eng = lb.FastWindingNumber(V, F, hierarchy="kd_sah", leaf_size=4)
    ref = oracle_mod.RefEngine(V, F)
    if cfg == 3:
        _, (o, s, d) = prim.config_queries(cfg, V, F)
        ins = eng.query_grid(o, s, d)[1]
        ins_ref, om_ref = ref.grid(o, s, d, want_omega=True)
    else:
        q = prim.near_surface_points(V, F, 1 << 20, seed=0xC0FFEE04) if cfg == 4 else prim.uniform_points_in_bbox(*prim.mesh_bbox(V), 1 << 21, seed=0xC0FFEE05)
        ins = eng.is_inside(q)
        om_ref = ref.solid_angle(q)
        ins_ref = ref.is_inside(q)
    w_ref = om_ref / FOUR_PI
    strict = band_mask(w_ref, 1e-3)
    mism = int((ins[strict] != ins_ref[strict]).sum())
    wide = band_mask(w_ref, 2e-2)
    print(f"cfg{cfg} kd_sah/leaf 4: strict-band mismatches vs the restatement {mism} of {int(strict.sum())}; outside |w-0.5|<=2e-2: "
          f"{int((ins[wide] != ins_ref[wide]).sum())}")
    assert np.array_equal(ins[wide], ins_ref[wide])
    assert mism <= 1e-4 * len(ins)
