#!/usr/bin/env python
"""CTA timeline of the tiled path: how busy the SMs are over one query_grid call, kernel by kernel.

The library records {start ns, end ns, SM id} per CTA of k_plan_block / k_tile_plan / k_tile_query when WN_TRACE_FILE is set
(diagnostics only; globaltimer has ~1 us resolution on some drivers, the numbers are for shares, not for absolute times).

    python tools/cta_timeline.py [--share R W] [--lanes L] [--run N]     # runs cfg2, prints a JSON summary

Output per launch: span (first start -> last end), busy = sum of CTA durations / (span x resident slots), the time at which
half / 90 % / 99 % of the CTAs had finished, and the idle tail (span - time when the slot occupancy drops below half).
"""
import argparse
import json
import os
import struct
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

TAGS = {1: "k_tile_plan", 2: "k_tile_query", 11: "k_plan_block L1", 12: "k_plan_block L2", 13: "k_plan_block L3"}


def read_trace(path):
    with open(path, "rb") as f:
        assert f.read(4) == b"WNTR"
        (nl,) = struct.unpack("<i", f.read(4))
        out = []
        for _ in range(nl):
            tag, lane = struct.unpack("<ii", f.read(8))
            (count,) = struct.unpack("<q", f.read(8))
            d = np.frombuffer(f.read(count * 32), dtype=np.uint64).reshape(count, 4)
            out.append((tag, lane, d))
    return out


def occupancy_profile(start, end, t0, t1, bins=200):
    """Mean number of CTAs in flight in each of `bins` slices of [t0, t1)."""
    edges = np.linspace(t0, t1, bins + 1)
    ev = np.concatenate([start, end])
    dv = np.concatenate([np.ones_like(start), -np.ones_like(end)])
    o = np.argsort(ev, kind="stable")
    ev, lvl = ev[o], np.cumsum(dv[o])
    # time-weighted level per bin
    prof = np.zeros(bins)
    idx = np.searchsorted(ev, edges)
    for b in range(bins):
        ts = np.concatenate([[edges[b]], ev[idx[b]:idx[b + 1]], [edges[b + 1]]])
        lv = np.concatenate([[lvl[idx[b] - 1] if idx[b] > 0 else 0], lvl[idx[b]:idx[b + 1]]])
        prof[b] = float(np.sum(np.diff(ts) * lv) / max(edges[b + 1] - edges[b], 1e-9))
    return prof


def summarise(launches):
    rows = []
    g0 = min(int(d[:, 0].min()) for _, _, d in launches)
    g1 = max(int(d[:, 1].max()) for _, _, d in launches)
    busy_total = 0.0
    for tag, lane, d in launches:
        st = d[:, 0].astype(np.float64) - g0
        en = d[:, 1].astype(np.float64) - g0
        dur = en - st
        span = en.max() - st.min()
        prof = occupancy_profile(st, en, st.min(), en.max())
        peak = prof.max()
        # the tail: from the moment the number of CTAs in flight falls below half its plateau for good
        below = np.nonzero(prof >= 0.5 * peak)[0]
        tail = span * (1.0 - (below.max() + 1) / len(prof)) if len(below) else 0.0
        ends = np.sort(en) - st.min()
        rows.append({
            "kernel": TAGS.get(tag, str(tag)), "lane": lane, "ctas": int(len(d)), "t0_us": round(st.min() / 1e3, 1), "span_us": round(span / 1e3, 1),
            "cta_us_mean": round(dur.mean() / 1e3, 2), "cta_us_max": round(dur.max() / 1e3, 2), "in_flight_peak": round(float(peak), 1),
            "in_flight_mean": round(float(dur.sum() / span), 1), "fill": round(float(dur.sum() / span / peak), 3),
            "t50_us": round(ends[len(ends) // 2] / 1e3, 1), "t99_us": round(ends[int(len(ends) * 0.99)] / 1e3, 1),
            "tail_below_half_us": round(tail / 1e3, 1), "sms": int(len(np.unique(d[:, 2])))})
        if tag == 2:
            # the query kernel in detail: which CTAs are the long ones, which finish last, and how evenly the SMs are loaded
            o = np.argsort(-dur)[:12]
            rows[-1]["longest"] = [[int(i), round(st[i] / 1e3 - rows[-1]["t0_us"], 1), round(dur[i] / 1e3, 1)] for i in o]
            o = np.argsort(-en)[:12]
            rows[-1]["last_to_finish"] = [[int(i), round(st[i] / 1e3 - rows[-1]["t0_us"], 1), round(dur[i] / 1e3, 1)] for i in o]
            sm = d[:, 2].astype(np.int64)
            per_sm = np.bincount(sm, weights=dur, minlength=int(sm.max()) + 1) / span
            rows[-1]["sm_load_min_mean_max"] = [round(float(per_sm[per_sm > 0].min()), 2), round(float(per_sm.mean()), 2), round(float(per_sm.max()), 2)]
            q = np.quantile(dur, [0.1, 0.5, 0.9, 0.99]) / 1e3
            rows[-1]["cta_us_quantiles_10_50_90_99"] = [round(float(x), 1) for x in q]
            dec = np.array_split(dur, 10)
            rows[-1]["cta_us_mean_by_launch_decile"] = [round(float(x.mean()) / 1e3, 1) for x in dec]
        busy_total += span
    return {"call_span_us": round((g1 - g0) / 1e3, 1), "sum_of_kernel_spans_us": round(busy_total / 1e3, 1), "launches": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--share", type=int, nargs=2, default=None, metavar=("RANK", "WORLD"), help="trace one rank's strided share of the lattice")
    ap.add_argument("--file", default=None, help="summarise an existing trace instead of running")
    a = ap.parse_args()
    if a.file:
        print(json.dumps(summarise(read_trace(a.file))))
        return
    import torch

    import lagrange_b200 as lb

    V, F = lb.primitive.generate_subdivided_sphere("icosahedron", 8)
    eng = lb.FastWindingNumber(V, F)
    n = 512
    o, s, d = np.full(3, -1.1, np.float32), np.full(3, 2.2 / n, np.float32), np.array([n, n, n], np.int64)
    out = torch.empty(n ** 3, dtype=torch.uint8, device="cuda")
    kw = dict(layers=tuple(a.share)) if a.share else {}
    for _ in range(3):
        eng.query_grid(o, s, d, out_inside=out, **kw)
    torch.cuda.synchronize()
    path = os.path.join(tempfile.mkdtemp(), "trace.bin")
    os.environ["WN_TRACE_FILE"] = path
    eng.query_grid(o, s, d, out_inside=out, **kw)
    torch.cuda.synchronize()
    del os.environ["WN_TRACE_FILE"]
    print(json.dumps(summarise(read_trace(path))))


if __name__ == "__main__":
    main()
