#!/usr/bin/env python
"""Experiment: top-down SAH split selection over the Morton-sorted triangle order (range boxes from a sparse table),
collapsed to a 4-ary hierarchy like the reference's builder does (expand the child with the largest area*count), imported
with wn_create_from_topology.   python tools/morton_sah_experiment.py [subdiv] [candidates]"""
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


def morton_order(V, F):
    c = V[F].mean(axis=1).astype(np.float64)
    lo = c.min(axis=0)
    ext = (c.max(axis=0) - lo).max()
    u = np.clip(((c - lo) / ext * (1 << 21)).astype(np.int64), 0, (1 << 21) - 1)

    def spread(x):
        x = x & 0x1FFFFF
        x = (x | (x << 32)) & 0x1F00000000FFFF
        x = (x | (x << 16)) & 0x1F0000FF0000FF
        x = (x | (x << 8)) & 0x100F00F00F00F00F
        x = (x | (x << 4)) & 0x10C30C30C30C30C3
        x = (x | (x << 2)) & 0x1249249249249249
        return x

    code = (spread(u[:, 0]) << 2) | (spread(u[:, 1]) << 1) | spread(u[:, 2])
    return np.argsort(code, kind="stable")


class RangeBoxes:
    def __init__(self, lo, hi):
        self.lo = [lo]
        self.hi = [hi]
        n = len(lo)
        k = 1
        while 2 * k <= n:
            pl, ph = self.lo[-1], self.hi[-1]
            self.lo.append(np.minimum(pl[:-k], pl[k:]))
            self.hi.append(np.maximum(ph[:-k], ph[k:]))
            k *= 2

    def area(self, i, j):
        """half surface area of the box of [i, j] inclusive (vectorised)."""
        n = j - i + 1
        k = np.floor(np.log2(n)).astype(np.int64)
        out = np.empty(len(i), dtype=np.float64)
        for kk in np.unique(k):
            m = k == kk
            a, b = i[m], j[m] - (1 << kk) + 1
            lo = np.minimum(self.lo[kk][a], self.lo[kk][b]).astype(np.float64)
            hi = np.maximum(self.hi[kk][a], self.hi[kk][b]).astype(np.float64)
            d = hi - lo
            out[m] = d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 2] * d[:, 0]
        return out


def build(V, F, C=16):
    order = morton_order(V, F)
    tv = V[F[order]]
    rb = RangeBoxes(tv.min(axis=1), tv.max(axis=1))
    N = len(F)
    # binary tree, level synchronous. node arrays grow per level.
    node_lo, node_hi = [np.array([0])], [np.array([N - 1])]
    left, right = [], []  # per level: child descriptors (as ranges); ids assigned later
    levels = []
    cur_lo, cur_hi = np.array([0], dtype=np.int64), np.array([N - 1], dtype=np.int64)
    all_lo, all_hi, all_split = [], [], []
    while len(cur_lo):
        n = cur_hi - cur_lo + 1
        split = np.full(len(cur_lo), -1, dtype=np.int64)
        m = n >= 2
        idx = np.flatnonzero(m)
        if len(idx):
            lo, hi, nn = cur_lo[idx], cur_hi[idx], n[idx]
            best_cost = np.full(len(idx), np.inf)
            best_k = lo.copy()
            for c in range(C):
                # candidate: last index of the left part
                k = lo + ((c + 1) * nn) // (C + 1) - 1
                k = np.clip(k, lo, hi - 1)
                cost = rb.area(lo, k) * (k - lo + 1) + rb.area(k + 1, hi) * (hi - k)
                better = cost < best_cost
                best_cost[better] = cost[better]
                best_k[better] = k[better]
            split[idx] = best_k
        all_lo.append(cur_lo)
        all_hi.append(cur_hi)
        all_split.append(split)
        if not len(idx):
            break
        nl_lo, nl_hi = cur_lo[idx], split[idx]
        nr_lo, nr_hi = split[idx] + 1, cur_hi[idx]
        cur_lo = np.concatenate([nl_lo, nr_lo])
        cur_hi = np.concatenate([nl_hi, nr_hi])
        keep = cur_hi > cur_lo  # ranges of one triangle are leaves: no node
        # keep order: children of node idx[q]: left = position q, right = position q + len(idx)
        levels.append((idx, keep))
        cur_lo, cur_hi = cur_lo[keep], cur_hi[keep]
    # assign ids: internal nodes = ranges with >= 2 triangles, numbered level by level
    offs = np.cumsum([0] + [len(a) for a in all_lo])
    total = offs[-1]
    child = np.full((total, 2), -1, dtype=np.int64)
    area_cnt = np.zeros(total)
    for L in range(len(all_lo)):
        lo, hi, split = all_lo[L], all_hi[L], all_split[L]
        area_cnt[offs[L]:offs[L + 1]] = rb.area(lo, hi) * (hi - lo + 1)
        if L >= len(levels):
            continue
        idx, keep = levels[L]
        q = len(idx)
        pos = np.cumsum(keep) - 1  # position of a kept child in the next level
        for side in range(2):
            clo = (lo[idx] if side == 0 else split[idx] + 1)
            chi = (split[idx] if side == 0 else hi[idx])
            single = chi == clo
            kept_pos = pos[side * q:(side + 1) * q]
            cid = np.where(single, -(order[clo] + 2), offs[L + 1] + kept_pos)
            child[offs[L] + idx, side] = cid
    return child, area_cnt


def collapse4(child, weight):
    new_id = {0: 0}
    rows = []
    queue = [0]
    qi = 0
    while qi < len(queue):
        u = queue[qi]
        qi += 1
        kids = [c for c in child[u] if c != -1]
        while len(kids) < 4:
            cand = [(weight[c], k) for k, c in enumerate(kids) if c >= 0]
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
                row.append(int(c))
        rows.append(row + [-1] * (4 - len(row)))
    return np.array(rows, dtype=np.int32)


def main():
    subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    V, F = prim.generate_subdivided_sphere("icosahedron", subdiv)
    t0 = time.perf_counter()
    child, w = build(V, F, C)
    t1 = time.perf_counter()
    topo4 = collapse4(child, w)
    t2 = time.perf_counter()
    n1 = 512
    origin, spacing, dims = (-1.1, -1.1, -1.1), (2.2 / n1,) * 3, (n1, n1, n1)
    n = n1**3
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    rep = {"build_s": t1 - t0, "collapse_s": t2 - t1, "binary_nodes": int(len(child)), "nodes4": int(len(topo4)), "candidates": C}
    b2 = np.where(child >= 0, child, np.where(child == -1, -1, child)).astype(np.int32)
    engines = {"morton_sah_binary": lb.FastWindingNumber(V, F, topology=b2), "morton_sah_4ary": lb.FastWindingNumber(V, F, topology=topo4)}
    for name, eng in engines.items():
        ms = timed(lambda: eng.query_grid(origin, spacing, dims, out_inside=out))
        ste = eng.query_stats_grid(origin, spacing, dims, tiling=True)
        rep[name] = {"ms_auto": ms, "Gq_s_auto": n / ms / 1e6,
                     "executed_tiled": {k: ste[k] / n for k in ("node_tests", "far_field_evals", "exact_triangles", "lane_slots")},
                     "entries": eng.info["num_entries"], "inside": int(out.sum().item())}
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
