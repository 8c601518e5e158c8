#!/usr/bin/env python
"""Experiment: PLOC (parallel locally-ordered clustering, Meister & Bittner 2018) over the Morton order, in numpy, imported
with wn_create_from_topology: what an agglomerative, surface-area driven hierarchy buys the query kernels (cfg2).
    python tools/ploc_experiment.py [subdiv] [radius]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lagrange_b200 as lb  # noqa: E402
from morton_sah_experiment import collapse4, morton_order, timed  # noqa: E402

prim = lb.primitive


def half_area(lo, hi):
    d = (hi - lo).astype(np.float64)
    return d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 2] * d[:, 0]


def ploc(V, F, radius=8):
    order = morton_order(V, F)
    tv = V[F[order]]
    lo, hi = tv.min(axis=1), tv.max(axis=1)
    N = len(F)
    cid = -(order.astype(np.int64) + 2)  # cluster -> child code (triangle) or internal node id (creation index for now)
    cnt = np.ones(N, dtype=np.int64)
    child = np.full((N - 1, 2), -1, dtype=np.int64)
    weight = np.zeros(N - 1)
    created = 0
    iters = 0
    while len(cid) > 1:
        n = len(cid)
        best = np.full(n, np.inf)
        nn = np.full(n, -1, dtype=np.int64)
        for d in range(1, min(radius, n - 1) + 1):
            a = half_area(np.minimum(lo[:-d], lo[d:]), np.maximum(hi[:-d], hi[d:]))
            # pair (i, i+d): candidate for i (forward) and for i+d (backward); ties -> smaller index wins
            m = a < best[:-d]
            best[:-d][m] = a[m]
            nn[:-d][m] = np.arange(n - d)[m] + d
            m = a < best[d:]
            best[d:][m] = a[m]
            nn[d:][m] = np.arange(n - d)[m]
        i = np.arange(n)
        mutual = (nn[nn] == i) & (i < nn)
        a_idx = np.flatnonzero(mutual)
        b_idx = nn[a_idx]
        k = len(a_idx)
        ids = created + np.arange(k)
        child[ids, 0] = cid[a_idx]
        child[ids, 1] = cid[b_idx]
        nlo = np.minimum(lo[a_idx], lo[b_idx])
        nhi = np.maximum(hi[a_idx], hi[b_idx])
        ncnt = cnt[a_idx] + cnt[b_idx]
        weight[ids] = half_area(nlo, nhi) * ncnt
        lo[a_idx], hi[a_idx], cnt[a_idx] = nlo, nhi, ncnt
        cid[a_idx] = ids
        keep = np.ones(n, dtype=bool)
        keep[b_idx] = False
        lo, hi, cnt, cid = lo[keep], hi[keep], cnt[keep], cid[keep]
        created += k
        iters += 1
    # renumber: root (last created) -> 0
    M = N - 1
    ren = lambda c: np.where(c >= 0, (M - 1) - c, c)
    out = np.empty_like(child)
    out[(M - 1) - np.arange(M)] = ren(child)
    w = np.empty_like(weight)
    w[(M - 1) - np.arange(M)] = weight
    return out, w, iters


def main():
    subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    radius = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    V, F = prim.generate_subdivided_sphere("icosahedron", subdiv)
    t0 = time.perf_counter()
    child, w, iters = ploc(V, F, radius)
    t1 = time.perf_counter()
    topo4 = collapse4(child, w)
    t2 = time.perf_counter()
    n1 = 512
    origin, spacing, dims = (-1.1, -1.1, -1.1), (2.2 / n1,) * 3, (n1, n1, n1)
    n = n1**3
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    rep = {"build_s": t1 - t0, "collapse_s": t2 - t1, "iterations": iters, "radius": radius, "nodes4": int(len(topo4))}
    engines = {"ploc_binary": lb.FastWindingNumber(V, F, topology=child.astype(np.int32)), "ploc_4ary": lb.FastWindingNumber(V, F, topology=topo4)}
    for name, eng in engines.items():
        ms = timed(lambda: eng.query_grid(origin, spacing, dims, out_inside=out))
        ste = eng.query_stats_grid(origin, spacing, dims, tiling=True)
        rep[name] = {"ms_auto": ms, "Gq_s_auto": n / ms / 1e6,
                     "executed_tiled": {k: ste[k] / n for k in ("node_tests", "far_field_evals", "exact_triangles", "lane_slots")},
                     "entries": eng.info["num_entries"], "inside": int(out.sum().item()), "max_depth": eng.info.get("max_depth")}
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
