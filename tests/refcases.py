"""Meshes that exercise every branch of the reference builder's split rules (shared by the CPU and GPU tiers)."""
import numpy as np


def _soup(rng, n, spread=1.0, size=0.05, clustered=None):
    c = rng.random((n, 3)) * spread
    if clustered is not None:
        # most triangles in one tiny cluster, a few far away: no balanced span boundary => order-statistic fallback
        c[: int(n * clustered)] *= 1e-3
    V = (c[:, None, :] + rng.normal(size=(n, 3, 3)) * size).reshape(-1, 3).astype(np.float32)
    return V, np.arange(3 * n, dtype=np.int32).reshape(n, 3)


def reference_builder_cases(prim, big=True):
    """name -> (V, F). Sizes around every threshold of the rule set (4 | 6 | 32 items), soups, clustered soups that force the
    fallback, coincident triangles (zero-extent boxes), point triangles, the reduced BASELINE configs."""
    rng = np.random.Generator(np.random.PCG64(7))
    cases = {}
    for n in list(range(1, 12)) + [31, 32, 33, 34, 40, 64, 65, 100, 257, 1000]:
        cases[f"soup{n}"] = _soup(rng, n)
    cases["clustered5000"] = _soup(rng, 5000, clustered=0.97)
    cases["clustered20000"] = _soup(rng, 20000, clustered=0.99, size=1e-4)
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    cases["identical500"] = (np.tile(tri, (500, 1)), np.arange(1500, dtype=np.int32).reshape(500, 3))
    cases["points100"] = (np.zeros((300, 3), np.float32), np.arange(300, dtype=np.int32).reshape(100, 3))
    # centres on a cubic curve along x: very uneven span occupancy at every level
    n = 3000
    x = (np.arange(n, dtype=np.float64) ** 3 * 1e-9).astype(np.float32)
    V = np.zeros((3 * n, 3), np.float32)
    V[:, 0] = np.repeat(x, 3)
    V[1::3, 1] = 1e-3
    V[2::3, 2] = 1e-3
    cases["cubic_line3000"] = (V, np.arange(3 * n, dtype=np.int32).reshape(n, 3))
    for cfg in (1, 2, 3, 4, 5):
        cases[f"cfg{cfg}_small"] = prim.config_mesh(cfg, small=True)
    if big:
        cases["torus_20k"] = prim.generate_torus(5.0, 1.0, 100, 50)
        cases["soup_167k"] = prim.make_soup(*prim.generate_torus(5.0, 1.0, 250, 200))
    return cases
