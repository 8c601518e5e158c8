#!/usr/bin/env python
"""The reference's own benchmark protocol (modules/winding/tests/test_fast_winding_number.cpp:85-125): 10 000 points drawn
in the mesh's bounding box from a default-seeded std::mt19937, one is_inside per point, count the hits. Run here once per
point (the reference's calling pattern) and once as a single batch.

    python tools/reference_benchmark.py [mesh.obj]      # default: procedural torus (dragon.obj is not in this image)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lagrange_b200 as lb  # noqa: E402
from lagrange_b200 import callers, io  # noqa: E402


def main():
    if len(sys.argv) > 1:
        mesh = io.load_obj(sys.argv[1], triangulate=True)
        V, F, name = mesh.vertices, mesh.facets.astype(np.int32), os.path.basename(sys.argv[1])
    else:
        V, F = lb.primitive.generate_torus(5, 1, 200, 100)
        name = "torus 200x100 (procedural)"
    lo, hi = V.min(axis=0), V.max(axis=0)
    n = 10000
    raw = callers.mt19937_uniform_float(3 * n, 0.0, 1.0).reshape(-1, 3)
    pts = (raw * (hi - lo)[None, :] + lo[None, :]).astype(np.float32)
    eng = lb.FastWindingNumber(V, F)
    eng.is_inside(pts[:16])
    t0 = time.perf_counter()
    hits = sum(int(eng.is_inside(p)) for p in pts)
    t1 = time.perf_counter()
    batch = eng.is_inside(pts)
    t2 = time.perf_counter()
    print(json.dumps({"mesh": name, "triangles": int(len(F)), "samples": n, "inside_single_calls": hits, "inside_batched": int(batch.sum()),
                      "single_calls_ms": 1e3 * (t1 - t0), "batched_ms": 1e3 * (t2 - t1), "build_ms": eng.info["build_ms"]}))


if __name__ == "__main__":
    main()
