"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck) runs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lagrange_b200 as lb  # noqa: E402

prim = lb.primitive
V, F = prim.generate_torus(5, 1, 40, 20)
os.environ["WN_TILE"] = "1"
for kw in ({}, {"leaf_size": 4}, {"hierarchy": "kd"}, {"hierarchy": "kd", "leaf_size": 4}, {"hierarchy": "kd_sah", "leaf_size": 4}):
    eng = lb.FastWindingNumber(V, F, **kw)
    o, s, d = prim.lattice_for_bbox(*prim.mesh_bbox(V), (50, 17, 45))
    a = eng.query_grid(o, s, d, want_omega=True)[0]
    b = eng.query_grid(o, s, d, want_omega=True, tiling=False)[0]
    q = prim.uniform_points_in_bbox(*prim.mesh_bbox(V), 5000, seed=1)
    c = eng.solid_angle(q)
    e = eng.solid_angle(q, tiling=False)
    x = eng.exact_solid_angle(q[:300])
    st = eng.query_stats_grid(o, s, d, tiling=True)
    sdf, active = eng.sdf_grid(o, s, d, 3.0 * float(s[0]))
    strided = eng.query_grid(o, s, d, layers=(1, 2))[1]
    print("grid tiled-vs-generic", float(np.abs(a - b).max()), "points", float(np.abs(c - e).max()), "exact", float(np.abs(x - e[:300]).max()), st)
print("sanitize_small done")
