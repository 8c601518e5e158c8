#!/usr/bin/env python
"""Experiment: top-down object-median splits along the longest axis of the centroid bounds (balanced k-d style hierarchy),
optionally choosing the split by SAH among a few positions around the median, collapsed to 4-ary, imported with
wn_create_from_topology.   python tools/median_experiment.py [subdiv] [mode: median|sah]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lagrange_b200 as lb  # noqa: E402
from morton_sah_experiment import collapse4, timed  # noqa: E402

prim = lb.primitive


def build_median(V, F, mode="median"):
    tv = V[F]
    cen = tv.mean(axis=1).astype(np.float32)
    blo, bhi = tv.min(axis=1), tv.max(axis=1)
    N = len(F)
    perm = np.arange(N)
    cur_lo, cur_hi = np.array([0], dtype=np.int64), np.array([N - 1], dtype=np.int64)
    all_lo, all_hi, all_split, levels = [], [], [], []
    perm_levels = []
    while len(cur_lo):
        n = cur_hi - cur_lo + 1
        idx = np.flatnonzero(n >= 2)
        split = np.full(len(cur_lo), -1, dtype=np.int64)
        if len(idx):
            lo, hi = cur_lo[idx], cur_hi[idx]
            # element -> active node index
            starts = lo
            lens = hi - lo + 1
            node_of = np.repeat(np.arange(len(idx)), lens)
            pos = np.concatenate([np.arange(a, b + 1) for a, b in zip(lo, hi)]) if len(idx) < 64 else (np.repeat(lo, lens) + (np.arange(lens.sum()) - np.repeat(np.cumsum(lens) - lens, lens)))
            elems = perm[pos]
            c = cen[elems]
            seg = np.cumsum(lens) - lens
            cmin = np.minimum.reduceat(c, seg, axis=0)
            cmax = np.maximum.reduceat(c, seg, axis=0)
            axis = np.argmax(cmax - cmin, axis=1)
            key = c[np.arange(len(elems)), axis[node_of]]
            o = np.lexsort((key, node_of))
            perm[pos] = elems[o]
            if mode == "median":
                split[idx] = lo + (lens // 2) - 1
            else:
                # SAH over 7 candidate positions (1/8 .. 7/8 of the sorted range) using exact prefix/suffix boxes per node:
                # computed with cumulative min/max via segmented scans (numpy: loop over candidates with reduceat)
                se = elems[o]
                best = np.full(len(idx), np.inf)
                bk = lo + lens // 2 - 1
                for cnum in range(1, 8):
                    k = np.clip((lens * cnum) // 8, 1, lens - 1)  # left count
                    # left part [seg, seg+k), right [seg+k, seg+len)
                    bounds = np.stack([seg, seg + k], axis=1).reshape(-1)
                    l_lo = np.minimum.reduceat(blo[se], bounds, axis=0)
                    l_hi = np.maximum.reduceat(bhi[se], bounds, axis=0)
                    La, Ra = (l_hi[0::2] - l_lo[0::2]).astype(np.float64), (l_hi[1::2] - l_lo[1::2]).astype(np.float64)
                    # note: reduceat segment for odd indices runs to the next even boundary = end of node: right part
                    cost = (La[:, 0] * La[:, 1] + La[:, 1] * La[:, 2] + La[:, 2] * La[:, 0]) * k + \
                           (Ra[:, 0] * Ra[:, 1] + Ra[:, 1] * Ra[:, 2] + Ra[:, 2] * Ra[:, 0]) * (lens - k)
                    m = cost < best
                    best[m] = cost[m]
                    bk[m] = (lo + k - 1)[m]
                split[idx] = bk
        all_lo.append(cur_lo)
        all_hi.append(cur_hi)
        all_split.append(split)
        if not len(idx):
            break
        nl_lo, nl_hi = cur_lo[idx], split[idx]
        nr_lo, nr_hi = split[idx] + 1, cur_hi[idx]
        nxt_lo = np.concatenate([nl_lo, nr_lo])
        nxt_hi = np.concatenate([nl_hi, nr_hi])
        keep = nxt_hi > nxt_lo
        levels.append((idx, keep))
        cur_lo, cur_hi = nxt_lo[keep], nxt_hi[keep]
    offs = np.cumsum([0] + [len(a) for a in all_lo])
    total = offs[-1]
    child = np.full((total, 2), -1, dtype=np.int64)
    weight = np.zeros(total)
    # final boxes per node for the collapse weight: area * count using final perm (ranges are contiguous in the final order)
    fl, fh = blo[perm], bhi[perm]
    for L in range(len(all_lo)):
        lo, hi, split = all_lo[L], all_hi[L], all_split[L]
        for q in range(0, len(lo), 1 << 16):
            a, b = lo[q:q + (1 << 16)], hi[q:q + (1 << 16)]
            if len(a) < 2048:
                w = []
                for x, y in zip(a, b):
                    d = (fh[x:y + 1].max(axis=0) - fl[x:y + 1].min(axis=0)).astype(np.float64)
                    w.append((d[0] * d[1] + d[1] * d[2] + d[2] * d[0]) * (y - x + 1))
                weight[offs[L] + q: offs[L] + q + len(a)] = w
            else:
                weight[offs[L] + q: offs[L] + q + len(a)] = (b - a + 1)  # deep levels: count is a fine proxy
        if L >= len(levels):
            continue
        idx, keep = levels[L]
        qn = len(idx)
        pos = np.cumsum(keep) - 1
        for side in range(2):
            clo = (lo[idx] if side == 0 else split[idx] + 1)
            chi = (split[idx] if side == 0 else hi[idx])
            single = chi == clo
            kept_pos = pos[side * qn:(side + 1) * qn]
            child[offs[L] + idx, side] = np.where(single, -(perm[clo] + 2), offs[L + 1] + kept_pos)
    return child, weight


def main():
    subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    mode = sys.argv[2] if len(sys.argv) > 2 else "median"
    V, F = prim.generate_subdivided_sphere("icosahedron", subdiv)
    t0 = time.perf_counter()
    child, w = build_median(V, F, mode)
    t1 = time.perf_counter()
    topo4 = collapse4(child, w)
    t2 = time.perf_counter()
    n1 = 512
    origin, spacing, dims = (-1.1, -1.1, -1.1), (2.2 / n1,) * 3, (n1, n1, n1)
    n = n1**3
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    rep = {"mode": mode, "build_s": t1 - t0, "collapse_s": t2 - t1, "nodes4": int(len(topo4))}
    engines = {"binary": lb.FastWindingNumber(V, F, topology=child.astype(np.int32)), "4ary": lb.FastWindingNumber(V, F, topology=topo4)}
    for name, eng in engines.items():
        ms = timed(lambda: eng.query_grid(origin, spacing, dims, out_inside=out))
        ste = eng.query_stats_grid(origin, spacing, dims, tiling=True)
        rep[name] = {"ms_auto": ms, "Gq_s_auto": n / ms / 1e6,
                     "executed_tiled": {k: ste[k] / n for k in ("node_tests", "far_field_evals", "exact_triangles", "lane_slots")},
                     "entries": eng.info["num_entries"], "inside": int(out.sum().item()), "max_depth": eng.info.get("max_depth")}
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
