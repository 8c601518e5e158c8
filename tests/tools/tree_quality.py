#!/usr/bin/env python
"""Query throughput of cfg2 on the GPU-built LBVH vs the restatement's 4-ary SAH topology (imported with
wn_create_from_topology): how much a higher-quality hierarchy would buy the same kernels.

    python tests/tools/tree_quality.py [subdiv] > gpurun_out/tree_quality.json
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import lagrange_b200 as lb  # noqa: E402
import oracle  # noqa: E402

prim = lb.primitive


def timed(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    V, F = prim.generate_subdivided_sphere("icosahedron", subdiv)
    n1 = 512
    origin, spacing, dims = (-1.1, -1.1, -1.1), (2.2 / n1,) * 3, (n1, n1, n1)
    n = n1**3
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    rep = {"triangles": int(len(F))}
    ref = oracle.RefEngine(V, F)
    engines = {"lbvh": lb.FastWindingNumber(V, F), "lbvh_leaf4": lb.FastWindingNumber(V, F, leaf_size=4),
               "sah4_oracle_topology": lb.FastWindingNumber(V, F, topology=ref.topology())}
    for name, eng in engines.items():
        ms = timed(lambda: eng.query_grid(origin, spacing, dims, out_inside=out))
        ms_g = timed(lambda: eng.query_grid(origin, spacing, dims, out_inside=out, tiling=False))
        st = eng.query_stats_grid(origin, spacing, dims)
        ste = eng.query_stats_grid(origin, spacing, dims, tiling=True)
        rep[name] = {"ms_auto": ms, "Gq_s_auto": n / ms / 1e6, "ms_generic": ms_g, "Gq_s_generic": n / ms_g / 1e6,
                     "per_point": {k: st[k] / n for k in ("node_tests", "far_field_evals", "exact_triangles")},
                     "executed_tiled": {k: ste[k] / n for k in ("node_tests", "far_field_evals", "exact_triangles", "lane_slots")},
                     "entries": eng.info["num_entries"], "inside": int(out.sum().item())}
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
