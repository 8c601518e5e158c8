"""CPU tier: the device build logic (lagrange_b200/csrc/wn_build_core.cuh, wn_device.cuh) run through the host
emulation harness (tests/emul) and compared with the oracle. Same source as the sm_100a kernels, sequential loops."""
import glob
import os

import numpy as np
import pytest

from conftest import FOUR_PI, small_config

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
TOL_OMEGA = 1e-4 * FOUR_PI  # BASELINE.json north_star: solid_angle within 1e-4 * 4 pi of the reference


def check_packed_structure(em, nT):
    rec, link, tris, order = em.packed()
    n = em.num_entries
    leaf = np.signbit(rec[0, :, 3])
    assert sorted(order.tolist()) == list(range(nT))  # triangle array is a permutation
    # internal skip links point forward and nest properly; leaves cover [0, nT) in order, each triangle once
    covered = 0
    for i in range(n):
        if leaf[i]:
            first, cnt = link[i] >> 4, (link[i] & 15) + 1
            assert first == covered
            covered += cnt
        else:
            assert i < link[i] <= n
            if i + 1 < n:
                assert link[i] > i + 1 or leaf[i]  # an internal entry has at least one child entry
    assert covered == nT
    # nesting: a child's subtree ends inside its parent's
    stack = []
    for i in range(n):
        while stack and stack[-1] <= i:
            stack.pop()
        end = i + 1 if leaf[i] else link[i]
        if stack:
            assert end <= stack[-1]
        stack.append(end)
    assert np.isinf(rec[0, 0, 3])  # root is never approximated
    return rec, link, tris, order


@pytest.mark.parametrize("cfg", [1, 2, 3, 5])
def test_imported_topology_moments_are_bit_identical_to_the_oracle(oracle_mod, emul_mod, prim, cfg):
    V, F, q, _ = small_config(prim, cfg)
    ref = oracle_mod.RefEngine(V, F)
    topo = ref.topology()
    em = emul_mod.EmulEngine(V, F, child=topo)
    assert em.error == 0 and em.num_internal == ref.num_nodes
    assert em.num_entries == ref.num_nodes + len(F)  # every node and every triangle has a record (A.4)
    bd, r23 = ref.boxdata(), em.ref23()
    nI = em.num_internal
    for i in range(nI):
        for s in range(4):
            c = topo[i, s]
            if c == -1:
                continue
            node = c if c >= 0 else nI - (c + 2)
            assert np.array_equal(bd[i, s], r23[node]), (i, s)
    check_packed_structure(em, len(F))


@pytest.mark.parametrize("cfg", [1, 2, 3, 5])
def test_emulated_traversal_matches_the_oracle_on_its_tree(oracle_mod, emul_mod, prim, cfg):
    V, F, q, _ = small_config(prim, cfg)
    q = q[:: max(1, len(q) // 3000)]
    ref = oracle_mod.RefEngine(V, F)
    em = emul_mod.EmulEngine(V, F, child=ref.topology())
    for beta in (2.0, 3.0):
        o_ref, c_ref = ref.solid_angle(q, beta=beta, counters=True)
        o_em, c_em = em.solid_angle(q, beta=beta, counters=True)
        assert np.abs(o_em - o_ref).max() < TOL_OMEGA
        # same accepted set per point: far-field evaluations and exact triangles agree exactly, lane tests too
        assert c_em[1] == c_ref[1] and c_em[2] == c_ref[2] and c_em[0] == c_ref[0], (c_em, c_ref)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_emulation_against_golden(emul_mod, path):
    g = np.load(path)
    em = emul_mod.EmulEngine(g["V"], g["F"], child=g["topology"])
    for beta in (2, 3):
        om, cnt = em.solid_angle(g["Q"], beta=float(beta), counters=True)
        assert np.abs(om - g[f"ref_beta{beta}"]).max() < TOL_OMEGA
        assert np.array_equal(cnt[1:], g[f"cnt_beta{beta}"][1:])


@pytest.mark.parametrize("leaf_size", [1, 4, 16])
@pytest.mark.parametrize("bits", [30, 63])
def test_lbvh_structure_and_accuracy(oracle_mod, emul_mod, prim, leaf_size, bits):
    V, F = prim.generate_torus(5, 1, 40, 20)
    em = emul_mod.EmulEngine(V, F, leaf_size=leaf_size, morton_bits=bits)
    assert em.error == 0 and em.width == 2 and em.num_internal == len(F) - 1
    rec, link, tris, order = check_packed_structure(em, len(F))
    # the triangle records are the mesh's triangles in depth-first order
    assert np.array_equal(tris[:, :, :3], V[F[order]])
    q = prim.uniform_points_in_bbox(*prim.mesh_bbox(V), 1500, seed=3)
    ex = oracle_mod.exact64(V, F, q)
    err = np.abs(em.solid_angle(q) - ex).max() / FOUR_PI
    assert err < 2e-2
    assert np.abs(em.solid_angle(q, beta=8.0) - ex).max() / FOUR_PI < 2e-4


def test_lbvh_moments_match_a_direct_merge(emul_mod, prim):
    """Root moments of the LBVH = moments of the whole mesh, whatever the tree: compare N and area-centroid."""
    V, F = prim.generate_subdivided_sphere("icosahedron", 3)
    em = emul_mod.EmulEngine(V, F)
    root = em.ref23(0, 1)[0]
    a, b, c = V[F[:, 0]].astype(np.float64), V[F[:, 1]].astype(np.float64), V[F[:, 2]].astype(np.float64)
    n = 0.5 * np.cross(b - a, c - a)
    assert np.abs(root[4:7] - n.sum(axis=0)).max() < 1e-5  # closed surface: ~0
    area = np.linalg.norm(n, axis=1)
    P = ((a + b + c) / 3 * area[:, None]).sum(axis=0) / area.sum()
    assert np.abs(root[0:3] - P).max() < 1e-5


def test_vertex_radius_is_never_larger_than_box_corner(emul_mod, prim):
    V, F = prim.generate_torus(5, 1, 30, 16)
    box = emul_mod.EmulEngine(V, F, radius_mode=0, approx_single=1).packed()[0]
    ver = emul_mod.EmulEngine(V, F, radius_mode=1, approx_single=1).packed()[0]
    r_box, r_ver = np.abs(box[0, 1:, 3]), np.abs(ver[0, 1:, 3])
    assert np.all(r_ver <= r_box) and np.mean(r_ver < r_box) > 0.5


def test_degenerate_inputs(oracle_mod, emul_mod):
    # zero-area and duplicate triangles, coincident centroids (Morton ties), a single triangle, two triangles
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [2, 2, 2]], dtype=np.float32)
    F = np.array([[0, 1, 2], [0, 1, 2], [0, 1, 1], [4, 4, 4], [0, 2, 3], [0, 3, 1], [1, 3, 2], [0, 2, 1]], dtype=np.int32)
    q = np.array([[0.1, 0.1, 0.1], [3, 3, 3], [0, 0, 0], [-1, 0.5, 0.2]], dtype=np.float32)
    ex = oracle_mod.exact64(V, F, q)
    for kw in ({}, {"leaf_size": 4}, {"morton_bits": 30}):
        em = emul_mod.EmulEngine(V, F, **kw)
        assert em.error == 0
        check_packed_structure(em, len(F))
        assert np.abs(em.solid_angle(q, beta=50.0) - ex).max() < 1e-4
    ref = oracle_mod.RefEngine(V, F)
    em = emul_mod.EmulEngine(V, F, child=ref.topology())
    assert np.abs(em.solid_angle(q) - ref.solid_angle(q)).max() < 1e-5
    for n in (1, 2):
        em = emul_mod.EmulEngine(V, F[4:4 + n])
        assert em.error == 0 and em.num_entries == 1 + n
        assert np.abs(em.solid_angle(q) - oracle_mod.exact64(V, F[4:4 + n], q)).max() < 1e-5


def test_malformed_topologies_are_rejected(emul_mod):
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float32)
    F = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int32)
    good = np.array([[-2, -3, -1, -1]], dtype=np.int32)
    assert emul_mod.EmulEngine(V, F, child=good).error == 0
    for bad in ([[-2, -2, -1, -1]],            # triangle 0 twice, triangle 1 missing
                [[-2, -1, -3, -1]],            # empty slot before a used one
                [[-2, -9, -1, -1]],            # triangle out of range
                [[1, -2, -1, -1], [0, -3, -1, -1]],   # child points at the root
                [[-2, -3, -1, -1], [-1, -1, -1, -1]]):  # unreferenced childless node
        assert emul_mod.EmulEngine(V, F, child=np.array(bad, dtype=np.int32)).error != 0, bad


def test_inside_threshold_matches_the_reference_expression(oracle_mod, emul_mod):
    L = emul_mod.lib()
    thr = np.float32(6.2831854820251465)
    x = np.nextafter(thr, np.float32(0))
    for _ in range(4):
        x = np.nextafter(x, np.float32(0))
    for _ in range(10):
        assert bool(L.emul_inside_from_omega(float(x))) == oracle_mod.inside_predicate(float(x))
        x = np.nextafter(x, np.float32(100))
    for v in (-1.0, 0.0, 6.28, 6.2832, 100.0, float("inf"), float("nan")):
        assert bool(L.emul_inside_from_omega(v)) == oracle_mod.inside_predicate(v)


@pytest.mark.parametrize("hierarchy", ["kd", "kd_sah"])
def test_kd_hierarchy_is_balanced_and_complete(emul_mod, prim, oracle_mod, hierarchy):
    """K3' / K3'' (wn_kd.cuh): the k-d hierarchies, emulated on the host from the device source."""
    V, F = prim.generate_torus(5, 1, 20, 11)  # 440 triangles: not a power of two
    em = emul_mod.EmulEngine(V, F, hierarchy=hierarchy)
    assert em.error == 0
    topo = em.topology()
    assert topo.shape == (len(F) - 1, 2)
    # every triangle exactly once, every internal node exactly once, depth = ceil(log2 n)
    seen_tri, seen_node = np.zeros(len(F), dtype=int), np.zeros(len(topo), dtype=int)
    depth = {0: 0}
    stack = [0]
    seen_node[0] = 1
    while stack:
        u = stack.pop()
        sizes = []
        for c in topo[u]:
            if c >= 0:
                seen_node[c] += 1
                depth[c] = depth[u] + 1
                stack.append(c)
            elif c <= -2:
                seen_tri[-(c + 2)] += 1
                depth[("t", -(c + 2))] = depth[u] + 1
    assert np.all(seen_tri == 1) and np.all(seen_node == 1)
    if hierarchy == "kd":
        assert max(depth.values()) == int(np.ceil(np.log2(len(F))))
    else:  # SAH cuts lie between 1/8 and 7/8 of a range
        assert max(depth.values()) <= int(np.ceil(np.log(len(F)) / np.log(8.0 / 7.0)))
    check_packed_structure(em, len(F))
    # same engine semantics on this hierarchy: the error against the exact winding number is the restatement's error class
    q = prim.uniform_points_in_bbox(*prim.mesh_bbox(V), 4000, seed=2)
    w = em.solid_angle(q) / (4 * np.pi)
    w_exact = oracle_mod.exact64(V, F, q) / (4 * np.pi)
    w_ref = oracle_mod.RefEngine(V, F).solid_angle(q) / (4 * np.pi)
    assert np.abs(w - w_exact).max() < 2.0 * max(np.abs(w_ref - w_exact).max(), 2e-3)
    assert np.abs(w - w_exact).mean() < 2.0 * np.abs(w_ref - w_exact).mean()


def _check_binary_topology(topo, nT):
    seen_tri, seen_node = np.zeros(nT, dtype=int), np.zeros(len(topo), dtype=int)
    stack = [0]
    seen_node[0] = 1
    while stack:
        u = stack.pop()
        for c in topo[u]:
            if c >= 0:
                seen_node[c] += 1
                stack.append(c)
            elif c <= -2:
                seen_tri[-(c + 2)] += 1
    assert np.all(seen_tri == 1) and np.all(seen_node == 1)


@pytest.mark.parametrize("hierarchy", ["kd", "kd_sah"])
def test_kd_builders_on_random_and_degenerate_soups(emul_mod, oracle_mod, hierarchy):
    """Random triangle soups of awkward sizes, with duplicated triangles, zero-area triangles and coincident centroids: the
    k-d builders must always produce a complete tree and the same winding numbers as the exact sum far from the soup."""
    rng = np.random.RandomState(7)
    for n in (2, 3, 5, 15, 16, 17, 31, 64, 127, 200):
        V = rng.randn(3 * n, 3).astype(np.float32)
        F = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
        if n >= 5:
            F[1] = F[0]                      # duplicate
            F[2] = [F[2, 0], F[2, 0], F[2, 1]]  # zero area
            V[F[3]] = V[F[4]]                # coincident copy (same centroid, same box)
        for leaf in (1, 4):
            em = emul_mod.EmulEngine(V, F, hierarchy=hierarchy, leaf_size=leaf)
            assert em.error == 0
            topo = em.topology()
            assert topo.shape == (n - 1, 2)
            _check_binary_topology(topo, n)
            check_packed_structure(em, n)
            q = (rng.randn(64, 3) * 30).astype(np.float32)  # far away: everything is expanded, errors are tiny
            w = em.solid_angle(q) / (4 * np.pi)
            w_exact = oracle_mod.exact64(V, F, q) / (4 * np.pi)
            assert np.abs(w - w_exact).max() < 2e-4
    # all triangles identical
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32)
    F = np.tile(np.array([[0, 1, 2]], dtype=np.int32), (40, 1))
    em = emul_mod.EmulEngine(V, F, hierarchy=hierarchy)
    assert em.error == 0
    _check_binary_topology(em.topology(), 40)
