"""CPU tier: K3R (WN_HIERARCHY_REFERENCE, lagrange_b200/csrc/wn_refbuild_core.cuh) run by the host emulation's sequential
backend must produce the restatement's UT_BVH<4> topology bit for bit — child table, depth-first numbering, order of the
triangle children (oracle/wn_oracle.cpp BvhBuilder). The GPU tier repeats this with the real kernels
(tests/test_gpu_reference_tree.py)."""
import numpy as np
import pytest

from refcases import reference_builder_cases


def test_reference_builder_emulation_equals_the_restatement(prim, oracle_mod, emul_mod):
    fallback_rounds = 0
    for name, (V, F) in reference_builder_cases(prim).items():
        topo, levels, syncs = emul_mod.ref_topology(V, F)
        fallback_rounds += emul_mod.ref_last_fallback_rounds()
        ref = oracle_mod.RefEngine(V, F).topology()
        assert topo.shape == ref.shape, (name, topo.shape, ref.shape)
        assert np.array_equal(topo, ref), (name, np.nonzero((topo != ref).any(1))[0][:5])
        assert levels <= 64 and syncs <= 4 * levels + 4, (name, levels, syncs)
    assert fallback_rounds > 0  # the order-statistic fallback was exercised


def test_reference_builder_structure(prim, emul_mod):
    """Structural facts of UT_BVH<4> with one item per leaf slot: every triangle exactly once, every node referenced once,
    nodes of more than four items have four children, children numbered depth first."""
    V, F = prim.generate_torus(5.0, 1.0, 60, 30)
    topo, _, _ = emul_mod.ref_topology(V, F)
    tris = -(topo[topo <= -2] + 2)
    assert np.array_equal(np.sort(tris), np.arange(len(F)))
    kids = topo[topo >= 0]
    assert np.array_equal(np.sort(kids), np.arange(1, len(topo)))
    # pre-order: the first node child of node i is i + 1; node children of one node increase
    for i, row in enumerate(topo):
        nk = row[row >= 0]
        if len(nk):
            assert nk[0] == i + 1 and np.all(np.diff(nk) > 0)
        assert not np.any((row[:-1] == -1) & (row[1:] != -1))  # empty slots are trailing


@pytest.mark.parametrize("n", [0])
def test_reference_builder_empty_mesh(emul_mod, n):
    topo, levels, _ = emul_mod.ref_topology(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32))
    assert topo.shape == (0, 4) and levels == 0
