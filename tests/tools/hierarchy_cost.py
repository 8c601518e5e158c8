#!/usr/bin/env python
"""Compare hierarchies without a GPU: the host emulation builds each tree from the device source and runs the cost model of
the tiled query path (emul_tile_cost) on a sample of BASELINE cfg2's tiles. Prints modelled warp-instructions per tile next
to the G queries/s measured on a B200 where that is known, so the model can be judged before it is trusted.

    python tests/tools/hierarchy_cost.py [subdiv=8] [tile_stride=4]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emul  # noqa: E402
import lagrange_b200 as lb  # noqa: E402

MEASURED = {("lbvh", 1): 5.12, ("lbvh", 4): 5.30, ("kd", 1): 5.75, ("kd", 4): 6.09, ("kd_sah", 4): 6.47}


def main():
    subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    stride = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    V, F = lb.primitive.generate_subdivided_sphere("icosahedron", subdiv)
    n1 = 512
    origin, spacing, dims = (-1.1, -1.1, -1.1), (2.2 / n1,) * 3, (n1, n1, n1)
    rows = []
    which = os.environ.get("HC_ONLY")
    combos = (("lbvh", 1), ("lbvh", 4), ("kd", 1), ("kd", 4), ("kd_sah", 1), ("kd_sah", 4))
    if which:
        combos = tuple((w.split(":")[0], int(w.split(":")[1])) for w in which.split(","))
    for h, leaf in combos:
        t0 = time.perf_counter()
        em = emul.EmulEngine(V, F, hierarchy=h, leaf_size=leaf)
        t1 = time.perf_counter()
        c = em.tile_cost(origin, spacing, dims, tile_stride=stride)
        t2 = time.perf_counter()
        c.update({"hierarchy": h, "leaf_size": leaf, "entries": em.num_entries, "build_s": t1 - t0, "model_s": t2 - t1,
                  "measured_Gq_s": MEASURED.get((h, leaf))})
        rows.append(c)
        print(json.dumps(c), flush=True)
    base = rows[0]["instr_per_tile"]
    for r in rows:
        m = r["measured_Gq_s"]
        print(f'{r["hierarchy"]:7s} leaf {r["leaf_size"]}: model {r["instr_per_tile"]:9.0f} instr/tile -> x{base / r["instr_per_tile"]:.3f} vs lbvh-1'
              + (f'   measured x{m / 5.12:.3f}' if m else ""))


if __name__ == "__main__":
    main()
