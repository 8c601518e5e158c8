"""Generates tests/golden/*.npz from the oracle (run from the repo root: python tests/golden/make_golden.py).

The reference holds no golden vectors for this path (parity unpinned), so these fixtures pin the
*restatement*: they freeze the oracle's outputs on small seeded inputs, so that (a) a later change to oracle/ that
alters results is noticed and (b) the GPU tests can compare against fixed numbers without re-deriving them.
Known-answer checks that do not depend on any implementation live in tests/test_oracle.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from lagrange_b200 import primitive as prim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.Generator(np.random.PCG64(20261017))
    cases = {
        "torus_24x12": prim.generate_torus(5.0, 1.0, 24, 12),
        "icosphere_l2": prim.generate_subdivided_sphere("icosahedron", 2),
        "soup_40x24": prim.make_soup(*prim.generate_torus(5.0, 1.0, 40, 24), seed=0xC0FFEE03),
    }
    for name, (V, F) in cases.items():
        lo, hi = prim.mesh_bbox(V)
        q = prim.uniform_points_in_bbox(lo, hi, 300, seed=int(rng.integers(1 << 30)))
        near = prim.near_surface_points(V, F, 200, sigma_rel=5e-3, seed=int(rng.integers(1 << 30)))
        q = np.concatenate([q, near], axis=0).astype(np.float32)
        ref = oracle.RefEngine(V, F)
        out = {"V": V, "F": F, "Q": q, "exact64": oracle.exact64(V, F, q), "topology": ref.topology()}
        for beta in (2.0, 3.0):
            om, cnt = ref.solid_angle(q, beta=beta, counters=True)
            out[f"ref_beta{int(beta)}"] = om
            out[f"cnt_beta{int(beta)}"] = cnt
        out["inside_beta2"] = ref.is_inside(q, beta=2.0)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, V.shape, F.shape, q.shape)


if __name__ == "__main__":
    main()
