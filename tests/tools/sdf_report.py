#!/usr/bin/env python
"""Throughput and full-size parity sample of the narrow-band signed distance (K10, wn_sdf_grid) on BASELINE cfg2's mesh:
1 310 720-triangle icosphere, 512^3 voxels over [-1.1, 1.1]^3, band = 3 voxels (what volume::mesh_to_volume asks of OpenVDB).

    python tests/tools/sdf_report.py > gpurun_out/sdf_report.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import lagrange_b200 as lb  # noqa: E402
import oracle  # noqa: E402  (checker only)

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
    V, F = prim.config_mesh(2)
    eng = lb.FastWindingNumber(torch.from_numpy(V).cuda(), torch.from_numpy(F).cuda())
    n1 = 512
    origin, spacing, dims = (-1.1, -1.1, -1.1), (2.2 / n1,) * 3, (n1, n1, n1)
    band = 3 * 2.2 / n1
    n = n1**3
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    ins = torch.empty(n, dtype=torch.uint8, device="cuda")
    rep = {"mesh_triangles": int(len(F)), "voxels": n, "band_voxels": 3}
    ms_signed = timed(lambda: eng.sdf_grid(origin, spacing, dims, band, out=out))
    sdf, active = eng.sdf_grid(origin, spacing, dims, band, out=out)
    ms_unsigned = timed(lambda: eng.sdf_grid(origin, spacing, dims, band, signed=False, out=out))
    ms_inside = timed(lambda: eng.query_grid(origin, spacing, dims, out_inside=ins))
    rep.update({"ms_signed": ms_signed, "ms_unsigned_distance_only": ms_unsigned, "ms_is_inside_only": ms_inside,
                "Gvoxels_per_s_signed": n / ms_signed / 1e6, "active_voxels": active, "active_fraction": active / n})
    # parity sample at full size: 2048 voxels drawn from the band and 2048 from everywhere, against the double brute force
    sdf, _ = eng.sdf_grid(origin, spacing, dims, band, out=out)
    h = sdf.cpu().numpy().reshape(-1)
    rng = np.random.RandomState(11)
    in_band = np.flatnonzero(np.abs(h) < np.float32(band))
    pick = np.concatenate([rng.choice(in_band, 2048, replace=False), rng.randint(0, n, 2048)])
    k, j, i = np.unravel_index(pick, (n1, n1, n1))
    sp = np.float32(2.2 / n1)
    q = np.stack([np.float32(-1.1) + sp * (i.astype(np.float32) + np.float32(0.5)), np.float32(-1.1) + sp * (j.astype(np.float32) + np.float32(0.5)),
                  np.float32(-1.1) + sp * (k.astype(np.float32) + np.float32(0.5))], axis=1).astype(np.float32)
    t0 = time.perf_counter()
    d = oracle.distance64(V, F, q)
    cpu_s = time.perf_counter() - t0
    r = np.linalg.norm(q.astype(np.float64), axis=1)
    err = np.abs(np.abs(h[pick]) - np.minimum(d, band))
    rep.update({"parity_sample": len(pick), "max_abs_err_vs_double_brute_force": float(err.max()),
                "sign_agrees_with_r_lt_1_outside_1e-3": bool(np.array_equal((h[pick] < 0)[np.abs(r - 1) > 1e-3], (r < 1)[np.abs(r - 1) > 1e-3])),
                "cpu_brute_force_s_for_sample": cpu_s, "cpu_threads": os.cpu_count()})
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
