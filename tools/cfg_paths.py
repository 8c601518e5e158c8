#!/usr/bin/env python
"""Generic vs forced-tiled path on one BASELINE config (device-resident), with the executed-work counters of each.

    python tools/cfg_paths.py 4 [reference|kd_sah|lbvh]
"""
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
    cfg = int(sys.argv[1])
    hierarchy = sys.argv[2] if len(sys.argv) > 2 else "reference"
    npts = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    V, F = prim.config_mesh(cfg)
    kind, q = prim.config_queries(cfg, V, F) if not npts or cfg not in (4, 5) else ("points", None)
    if q is None:
        q = prim.near_surface_points(V, F, npts, seed=0xC0FFEE04) if cfg == 4 else prim.uniform_points_in_bbox(*prim.mesh_bbox(V), npts, seed=0xC0FFEE05)
    eng = lb.FastWindingNumber(V, F, hierarchy=hierarchy, leaf_size=4)
    out = {"cfg": cfg, "hierarchy": hierarchy, "triangles": int(len(F)), "build_ms": eng.info["build_ms"], "entries": eng.info["num_entries"]}
    if kind == "grid":
        o, s, d = q
        n = int(np.prod(d))
        ob = torch.empty(n, dtype=torch.uint8, device="cuda")
        run = lambda tiling: eng.query_grid(o, s, d, out_inside=ob, tiling=tiling)
        stats = lambda tiling: eng.query_stats_grid(o, s, d, tiling=tiling)
    else:
        n = len(q)
        dq = torch.from_numpy(q).cuda()
        ob = torch.empty(n, dtype=torch.uint8, device="cuda")
        run = lambda tiling: eng.is_inside(dq, out=ob, tiling=tiling)
        stats = lambda tiling: eng.query_stats(dq, tiling=tiling)
        t_sorted = None
    out["n"] = n
    for name, env, tiling in (("auto", None, True), ("generic", None, False), ("forced_tiled", "1", True)):
        if env is None:
            os.environ.pop("WN_TILE", None)
        else:
            os.environ["WN_TILE"] = env
        ms = timed(lambda: run(tiling))
        st = stats(tiling)
        out[name] = {"ms": ms, "gqps": n / ms / 1e6, "tests": st["node_tests"] / n, "evals": st["far_field_evals"] / n, "exact": st["exact_triangles"] / n,
                     "lane_util": st["node_tests"] / max(1, st["lane_slots"]), "inside": int(ob.sum().item())}
    os.environ.pop("WN_TILE", None)
    if kind != "grid":
        # how much of the point path is the Morton sort of the queries (K9)?
        ms_pre = timed(lambda: eng.is_inside(dq, out=ob, tiling=False, presorted=True))
        out["generic_presorted_flag_unsorted_points"] = {"ms": ms_pre}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
