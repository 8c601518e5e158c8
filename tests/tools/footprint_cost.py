#!/usr/bin/env python
"""Cost-model sweep (no GPU) over what k_tile_query could do differently on the final hierarchy: the warp's footprint, the
number of queries per lane, the tile size. Prints modelled warp-instructions per lattice point (query + plan).
Set WN_EMUL_WIDE=0 to model the binary packing instead of the 4-ary one.

    python tests/tools/footprint_cost.py [subdiv=8]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emul  # noqa: E402
import lagrange_b200 as lb  # noqa: E402

CASES = (  # tile shape, points per warp, tile stride of the sample
    ((8, 8, 8), (4, 4, 4), 4),    # what the kernels do
    ((8, 8, 8), (8, 4, 2), 4), ((8, 8, 8), (8, 8, 1), 4), ((8, 8, 8), (2, 4, 8), 4),
    ((8, 8, 8), (4, 4, 2), 4),    # one query per lane
    ((8, 8, 8), (4, 4, 8), 4),    # four queries per lane
    ((8, 8, 16), (4, 4, 8), 4), ((16, 8, 8), (4, 4, 8), 4), ((16, 16, 8), (4, 4, 8), 2), ((8, 8, 4), (4, 4, 4), 4),
)


def main():
    subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    V, F = lb.primitive.generate_subdivided_sphere("icosahedron", subdiv)
    n1 = 512
    origin, spacing, dims = (-1.1,) * 3, (2.2 / n1,) * 3, (n1,) * 3
    em = emul.EmulEngine(V, F, hierarchy="kd_sah", leaf_size=4)
    for tile, warp, stride in CASES:
        c = em.tile_cost(origin, spacing, dims, tile_stride=stride, warp_shape=warp, tile_shape=tile)
        pts = float(np.prod(tile))
        print(f"tile {tile} warp {warp}: {c['instr_per_point']:6.1f} instr/point (query {c['query_instr'] / pts:5.1f}, plan {c['plan_instr'] / pts:5.1f}); "
              f"far set {c['far_set']:.1f}, conditional items {c['conditional_items']:.1f}, walk steps {c['walk_steps']:.1f}", flush=True)


if __name__ == "__main__":
    main()
