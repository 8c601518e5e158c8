"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck) runs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lagrange_b200 as lb  # noqa: E402

prim = lb.primitive
V, F = prim.generate_torus(5, 1, 40, 20)
os.environ["WN_TILE"] = "1"
os.environ["WN_TILE_SPLIT_MIN"] = "64"  # host outputs: exercise the batch split + overlapped copies on a small lattice
for kw in ({"hierarchy": "reference"}, {"hierarchy": "lbvh"}, {"hierarchy": "lbvh", "leaf_size": 4}, {"hierarchy": "kd"}, {"hierarchy": "kd", "leaf_size": 4},
           {"hierarchy": "kd_sah", "leaf_size": 4}):
    eng = lb.FastWindingNumber(V, F, **kw)
    o, s, d = prim.lattice_for_bbox(*prim.mesh_bbox(V), (50, 40, 45))  # 7 x 5 x 6 tiles: hierarchical planning blocks with ragged edges
    a = eng.query_grid(o, s, d, want_omega=True)[0]
    b = eng.query_grid(o, s, d, want_omega=True, tiling=False)[0]
    bits = eng.query_grid(o, s, d, bits=True)[1]
    q = prim.uniform_points_in_bbox(*prim.mesh_bbox(V), 5000, seed=1)
    c = eng.solid_angle(q)
    e = eng.solid_angle(q, tiling=False)
    x = eng.exact_solid_angle(q[:300])
    x2 = eng.exact_solid_angle(q[:3000])
    st = eng.query_stats_grid(o, s, d, tiling=True)
    sdf, active = eng.sdf_grid(o, s, d, 3.0 * float(s[0]))
    idx, val, sbits = eng.sdf_grid_sparse(o, s, d, 3.0 * float(s[0]), want_inside_bits=True)
    sq, tri, xyz = eng.closest_point(q)
    strided = eng.query_grid(o, s, d, layers=(1, 2))[1]
    strided_bits = eng.query_grid(o, s, d, layers=(0, 3), bits=True)[1]
    o2, s2, d2 = prim.lattice_for_bbox(*prim.mesh_bbox(V), (24, 64, 40))  # ny % 32 == 0: the diagonal sharding splits rows in 4 / 2 parts
    for rank, world in ((0, 4), (3, 4), (1, 2), (2, 8)):
        eng.query_grid(o2, s2, d2, shard=(rank, world), bits=(rank == 3))
    if kw["hierarchy"] == "reference":
        # the CTA timeline hooks of the tiled kernels (diagnostics)
        import tempfile

        os.environ["WN_TRACE_FILE"] = os.path.join(tempfile.mkdtemp(), "trace.bin")
        eng.query_grid(o, s, d)
        del os.environ["WN_TRACE_FILE"]
    rep = lb.FastWindingNumber.from_packed(eng.pack())
    assert np.array_equal(rep.solid_angle(q[:100], tiling=False), e[:100])
    print(kw, "grid tiled-vs-generic", float(np.abs(a - b).max()), "points", float(np.abs(c - e).max()), "exact", float(np.abs(x - e[:300]).max()),
          "band", len(idx), active, st)
# the reference builder's order-statistic fallback (clustered soup) and its tiny-mesh paths
rng = np.random.Generator(np.random.PCG64(7))
c = rng.random((3000, 3))
c[:2900] *= 1e-3
Vs = (c[:, None, :] + rng.normal(size=(3000, 3, 3)) * 1e-4).reshape(-1, 3).astype(np.float32)
Fs = np.arange(9000, dtype=np.int32).reshape(3000, 3)
for n in (1, 3, 5, 33, 3000):
    lb.FastWindingNumber(Vs, Fs[:n], hierarchy="reference").solid_angle(Vs[:50] + 0.01)
print("sanitize_small done")
