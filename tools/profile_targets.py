#!/usr/bin/env python
"""One launch each of the kernels that had no ncu summary yet: the generic k_query on a coherent
lattice, k_exact, k_sdf_grid. Run under ncu:

    ncu --set full --clock-control none -k regex:'k_query|k_exact|k_sdf_grid' -c 8 -o gpurun_out/r2_targets python tools/profile_targets.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lagrange_b200 as lb  # noqa: E402

prim = lb.primitive


def main():
    hierarchy = os.environ.get("WN_PROFILE_HIERARCHY", "kd_sah")
    V, F = prim.generate_subdivided_sphere("icosahedron", 7)  # 327 680 triangles
    eng = lb.FastWindingNumber(V, F, hierarchy=hierarchy, leaf_size=4)
    n = 256
    o, s, d = np.full(3, -1.1, np.float32), np.full(3, 2.2 / n, np.float32), np.array([n, n, n], np.int64)
    out = torch.empty(n ** 3, dtype=torch.uint8, device="cuda")
    eng.query_grid(o, s, d, out_inside=out, tiling=False)  # k_query<2, GRID>
    sdf = torch.empty(n ** 3, dtype=torch.float32, device="cuda")
    eng.sdf_grid(o, s, d, 3 * 2.2 / n, signed=False, out=sdf)  # k_sdf_grid
    V5, F5 = prim.config_mesh(5)
    e5 = lb.FastWindingNumber(V5, F5)
    q = torch.from_numpy(prim.uniform_points_in_bbox(*prim.mesh_bbox(V5), 1 << 16)).cuda()
    om = torch.empty(len(q), dtype=torch.float32, device="cuda")
    e5.exact_solid_angle(q, out=om)  # k_exact
    torch.cuda.synchronize()
    print("ok", int(out.sum().item()), float(sdf.min().item()), float(om.abs().max().item()))


if __name__ == "__main__":
    main()
