"""CPU tier: the diagnostics tooling works without a GPU (the CTA timeline reader on a synthetic trace in the library's dump format)."""
import importlib.util
import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_cta_timeline_reads_and_summarises_a_trace(tmp_path):
    tl = _load("cta_timeline")
    # two launches: a plan kernel of 100 CTAs (10 us each, 50 in flight) and a query kernel of 20 CTAs, 5 slots, one long CTA first
    t0 = 1_000_000
    plan = np.zeros((100, 4), np.uint64)
    for i in range(100):
        plan[i] = (t0 + (i // 50) * 10_000, t0 + (i // 50) * 10_000 + 10_000, i % 8, 0)
    q0 = t0 + 25_000
    query = np.zeros((20, 4), np.uint64)
    query[0] = (q0, q0 + 400_000, 0, 0)
    for i in range(1, 20):
        s = q0 + ((i - 1) // 4) * 80_000
        query[i] = (s, s + 80_000, i % 8, 0)
    path = tmp_path / "trace.bin"
    with open(path, "wb") as f:
        f.write(b"WNTR" + struct.pack("<i", 2))
        for tag, lane, d in ((1, 0, plan), (2, 0, query)):
            f.write(struct.pack("<iiq", tag, lane, len(d)) + d.tobytes())
    launches = tl.read_trace(str(path))
    assert [(t, ln, len(d)) for t, ln, d in launches] == [(1, 0, 100), (2, 0, 20)]
    s = tl.summarise(launches)
    p, q = s["launches"]
    assert p["kernel"] == "k_tile_plan" and p["ctas"] == 100 and abs(p["span_us"] - 20.0) < 1e-6 and abs(p["in_flight_mean"] - 50.0) < 1e-6
    assert q["kernel"] == "k_tile_query" and q["ctas"] == 20 and abs(q["cta_us_max"] - 400.0) < 1e-6
    assert q["longest"][0][0] == 0 and q["sms"] == 8
    assert abs(s["call_span_us"] - 425.0) < 1e-6
