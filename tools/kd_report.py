#!/usr/bin/env python
"""LBVH vs balanced k-d hierarchy on every BASELINE config: build ms and query rate on one GPU (device-resident).
    python tools/kd_report.py [configs] > gpurun_out/kd_report.json"""
import json
import os
import sys

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


LEAVES = tuple(int(x) for x in os.environ.get("KD_LEAVES", "1,4").split(","))


def main():
    cfgs = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "1,2,3,4,5").split(",")]
    rep = {}
    lb.FastWindingNumber(*prim.generate_subdivided_sphere("icosahedron", 4), hierarchy="kd").close()
    for cfg in cfgs:
        V, F = prim.config_mesh(cfg)
        kind, q = prim.config_queries(cfg, V, F)
        dV, dF = torch.from_numpy(V).cuda(), torch.from_numpy(F).cuda()
        row = {"triangles": int(len(F))}
        for h in ("lbvh", "kd"):
            for leaf in LEAVES:
                lb.FastWindingNumber(dV, dF, hierarchy=h, leaf_size=leaf).close()  # warm
                eng = lb.FastWindingNumber(dV, dF, hierarchy=h, leaf_size=leaf)
                info = eng.info
                if kind == "grid":
                    o, s, d = q
                    n = int(np.prod(d))
                    out = torch.empty(n, dtype=torch.uint8, device="cuda")
                    ms = timed(lambda: eng.query_grid(o, s, d, out_inside=out))
                else:
                    dq = torch.from_numpy(q).cuda()
                    n = len(q)
                    out = torch.empty(n, dtype=torch.uint8, device="cuda")
                    ms = timed(lambda: eng.is_inside(dq, out=out))
                row[f"{h}_leaf{leaf}"] = {"build_ms": info["build_ms"], "query_ms": ms, "Gq_s": n / ms / 1e6, "inside": int(out.sum().item()),
                                          "tree_mb": info["tree_bytes"] / 1e6, "max_depth": info["max_depth"]}
                eng.close()
        rep[f"cfg{cfg}"] = row
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
