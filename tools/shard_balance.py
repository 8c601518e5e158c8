#!/usr/bin/env python
"""Per-rank time of one cfg2 step under the two lattice shardings, measured on ONE GPU (rank by rank): shows how much of the
multi-GPU efficiency loss is imbalance between ranks and how much is fixed cost per call.

    python tools/shard_balance.py [world]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lagrange_b200 as lb  # noqa: E402


def timed(fn, reps=5):
    fn()
    fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    V, F = lb.primitive.generate_subdivided_sphere("icosahedron", 8)
    eng = lb.FastWindingNumber(V, F)
    n = 512
    o, s, d = np.full(3, -1.1, np.float32), np.full(3, 2.2 / n, np.float32), np.array([n, n, n], np.int64)
    out = torch.empty(n ** 3, dtype=torch.uint8, device="cuda")
    full = timed(lambda: eng.query_grid(o, s, d, out_inside=out))
    res = {"world": world, "full_ms": full, "ideal_ms": full / world}
    for name, kw in (("diagonal", lambda r: dict(shard=(r, world))), ("layers", lambda r: dict(layers=(r, world)))):
        ts = [timed(lambda: eng.query_grid(o, s, d, out_inside=out, **kw(r))) for r in range(world)]
        res[name] = {"per_rank_ms": [round(t, 4) for t in ts], "max_ms": max(ts), "mean_ms": float(np.mean(ts)), "sum_ms": float(np.sum(ts)),
                     "efficiency_if_parallel": full / world / max(ts)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
