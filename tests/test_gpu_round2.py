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
