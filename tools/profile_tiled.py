#!/usr/bin/env python
"""One tiled batch of cfg2 (k_tile_plan + k_tile_query) for ncu:

    ncu --set full --clock-control none --import-source on -k regex:'k_tile' -c 2 -f -o gpurun_out/r2_tiled python tools/profile_tiled.py [grid]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lagrange_b200 as lb  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    hierarchy = os.environ.get("WN_PROFILE_HIERARCHY", "reference")
    V, F = lb.primitive.generate_subdivided_sphere("icosahedron", 8)
    eng = lb.FastWindingNumber(V, F, hierarchy=hierarchy, leaf_size=4)
    o, s, d = np.full(3, -1.1, np.float32), np.full(3, 2.2 / n, np.float32), np.array([n, n, n], np.int64)
    out = torch.empty(n ** 3, dtype=torch.uint8, device="cuda")
    eng.query_grid(o, s, d, out_inside=out)
    torch.cuda.synchronize()
    print("inside", int(out.sum().item()))


if __name__ == "__main__":
    main()
