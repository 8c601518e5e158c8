#!/usr/bin/env python
"""Experiment: collapse the GPU-built binary LBVH into a 4-ary hierarchy on the host and re-import it, to see what a wide
tree buys the query kernels (cfg2).   python tools/collapse_experiment.py [subdiv]"""
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


def subtree_counts(child):
    n = len(child)
    cnt = np.zeros(n, dtype=np.int64)
    # children have larger depth; process nodes in reverse BFS order
    order = [0]
    i = 0
    while i < len(order):
        u = order[i]
        for c in child[u]:
            if c >= 0:
                order.append(c)
        i += 1
    for u in reversed(order):
        s = 0
        for c in child[u]:
            s += cnt[c] if c >= 0 else (1 if c <= -2 else 0)
        cnt[u] = s
    return cnt


def collapse(child, mode, cnt=None):
    """child: (n, 2) neutral encoding. Returns (m, 4) table."""
    new_id = {0: 0}
    rows = []
    queue = [0]
    qi = 0
    while qi < len(queue):
        u = queue[qi]
        qi += 1
        kids = [c for c in child[u] if c != -1]
        if mode == "grandchildren":
            out = []
            for c in kids:
                if c >= 0:
                    out.extend([g for g in child[c] if g != -1])
                else:
                    out.append(c)
            kids = out
        else:  # adaptive: expand the internal child with the most triangles until 4
            while len(kids) < 4:
                cand = [(cnt[c], k) for k, c in enumerate(kids) if c >= 0 and len(kids) - 1 + sum(1 for g in child[c] if g != -1) <= 4]
                if not cand:
                    break
                _, k = max(cand)
                c = kids.pop(k)
                kids.extend([g for g in child[c] if g != -1])
        row = []
        for c in kids:
            if c >= 0:
                new_id[c] = len(queue)
                queue.append(c)
                row.append(new_id[c])
            else:
                row.append(c)
        rows.append(row + [-1] * (4 - len(row)))
    return np.array(rows, dtype=np.int32)


def main():
    subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    V, F = prim.generate_subdivided_sphere("icosahedron", subdiv)
    n1 = 512
    origin, spacing, dims = (-1.1, -1.1, -1.1), (2.2 / n1,) * 3, (n1, n1, n1)
    n = n1**3
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    base = lb.FastWindingNumber(V, F, keep_build_data=True)
    topo = base.debug_topology()
    cnt = subtree_counts(topo)
    rep = {}
    engines = {"lbvh": base, "grandchildren": lb.FastWindingNumber(V, F, topology=collapse(topo, "grandchildren")),
               "adaptive_count": lb.FastWindingNumber(V, F, topology=collapse(topo, "adaptive", cnt))}
    for name, eng in engines.items():
        ms = timed(lambda: eng.query_grid(origin, spacing, dims, out_inside=out))
        st = eng.query_stats_grid(origin, spacing, dims)
        ste = eng.query_stats_grid(origin, spacing, dims, tiling=True)
        rep[name] = {"ms_auto": ms, "Gq_s_auto": n / ms / 1e6,
                     "per_point": {k: st[k] / n for k in ("node_tests", "far_field_evals", "exact_triangles")},
                     "executed_tiled": {k: ste[k] / n for k in ("node_tests", "far_field_evals", "exact_triangles", "lane_slots")},
                     "entries": eng.info["num_entries"], "inside": int(out.sum().item())}
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
