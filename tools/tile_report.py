#!/usr/bin/env python
"""Per-tile class sizes (probe) and executed-work counters of the tiled path vs the generic path for one BASELINE config.

    WN_VERBOSE=1 python tools/tile_report.py [cfg]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lagrange_b200 as lb  # noqa: E402

prim = lb.primitive
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
V, F = prim.config_mesh(cfg)
kind, q = prim.config_queries(cfg, V, F)
eng = lb.FastWindingNumber(torch.from_numpy(V).cuda(), torch.from_numpy(F).cuda())
assert kind == "grid"
o, s, d = q
n = int(np.prod(d))
out = torch.empty(n, dtype=torch.uint8, device="cuda")
eng.query_grid(o, s, d, out_inside=out)
rep = {}
for name, tiling in (("generic", False), ("tiled", True)):
    st = eng.query_stats_grid(o, s, d, tiling=tiling)
    rep[name] = {k: st[k] / n for k in ("node_tests", "far_field_evals", "exact_triangles", "lane_slots")}
print(json.dumps(rep))
hdr = eng.debug_last_plan()
if len(hdr):
    dist = {}
    for j, name in enumerate(("n_cond", "n_dir", "n_tri")):
        v = hdr[:, j]
        dist[name] = {"mean": float(v.mean()), "p50": int(np.percentile(v, 50)), "p90": int(np.percentile(v, 90)), "p99": int(np.percentile(v, 99)),
                      "max": int(v.max())}
    dist["fallback_tiles"] = int((hdr[:, 3] & 1).sum())
    dist["tiles"] = int(len(hdr))
    w = hdr[:, 0].astype(np.float64)
    order = np.sort(w)[::-1]
    dist["share_of_cond_in_top10pct_tiles"] = float(order[: len(order) // 10].sum() / max(w.sum(), 1))
    print(json.dumps({"last_batch_tiles": dist}))
