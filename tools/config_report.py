#!/usr/bin/env python
"""Throughput / build-time report for all five BASELINE.json configs on one GPU (device-resident inputs and outputs).
cfg2 is the bench.py headline; the others are parity-test cases whose rates are recorded here for DESIGN.md.

    python tools/config_report.py [--configs 1,3,4,5] > gpurun_out/config_report.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lagrange_b200 as lb  # noqa: E402

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
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    args = ap.parse_args()
    out = {}
    lb.FastWindingNumber(*prim.generate_subdivided_sphere("icosahedron", 4)).close()  # load the build kernels
    for cfg in [int(c) for c in args.configs.split(",")]:
        V, F = prim.config_mesh(cfg)
        kind, q = prim.config_queries(cfg, V, F)
        dV, dF = torch.from_numpy(V).cuda(), torch.from_numpy(F).cuda()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng = lb.FastWindingNumber(dV, dF)
        torch.cuda.synchronize()
        wall = 1e3 * (time.perf_counter() - t0)
        info = eng.info
        rec = {"triangles": int(len(F)), "build_ms": info["build_ms"], "build_wall_ms": wall, "tree_mb": info["tree_bytes"] / 1e6,
               "build_stages_ms": {k: info[k] for k in ("build_ms_morton", "build_ms_sort", "build_ms_hierarchy", "build_ms_moments", "build_ms_pack")}}
        if kind == "grid":
            o, s, d = q
            n = int(np.prod(d))
            outb = torch.empty(n, dtype=torch.uint8, device="cuda")
            ms = timed(lambda: eng.query_grid(o, s, d, out_inside=outb))
            ms_generic = timed(lambda: eng.query_grid(o, s, d, out_inside=outb, tiling=False))
            st = eng.query_stats_grid(o, s, d)
            ste = eng.query_stats_grid(o, s, d, tiling=True)
            rec["queries"] = f"{d[0]}x{d[1]}x{d[2]} lattice"
        else:
            n = len(q)
            dq = torch.from_numpy(q).cuda()
            outb = torch.empty(n, dtype=torch.uint8, device="cuda")
            ms = timed(lambda: eng.is_inside(dq, out=outb))  # includes the Morton sort of the queries (K9)
            ms_generic = timed(lambda: eng.is_inside(dq, out=outb, tiling=False))
            st = eng.query_stats(dq)
            ste = eng.query_stats(dq, tiling=True)
            rec["queries"] = f"{n} points"
        rec.update({"n_queries": n, "ms": ms, "gqps": n / ms / 1e6, "ms_generic": ms_generic, "gqps_generic": n / ms_generic / 1e6,
                    "per_point": {k: st[k] / n for k in ("node_tests", "far_field_evals", "exact_triangles")},
                    "executed": {k: ste[k] / n for k in ("node_tests", "far_field_evals", "exact_triangles")}})
        if cfg == 5:
            m = 1 << 20
            dq1 = dq[:m].contiguous()
            oe = torch.empty(m, dtype=torch.float32, device="cuda")
            ms_e = timed(lambda: eng.exact_solid_angle(dq1, out=oe), reps=2)
            rec["exact_mode"] = {"queries": m, "ms": ms_e, "gpairs_per_s": m * len(F) / ms_e / 1e6, "tflops_75_per_pair": 75 * m * len(F) / ms_e / 1e9}
        out[f"cfg{cfg}"] = rec
        print(f"cfg{cfg}", json.dumps(rec), file=sys.stderr, flush=True)
        del eng
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
