import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)

FOUR_PI = 4.0 * np.pi


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def prim():
    from lagrange_b200 import primitive

    return primitive


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def emul_mod():
    import emul

    emul.build()
    return emul


def small_config(prim, cfg):
    """(V, F, queries float32 [n,3], lattice or None) for the reduced-size version of BASELINE config `cfg`."""
    V, F = prim.config_mesh(cfg, small=True)
    kind, q = prim.config_queries(cfg, V, F, small=True)
    if kind == "grid":
        return V, F, prim.lattice_points(*q), q
    return V, F, q, None


def band_mask(w_ref, band=1e-3):
    """Points whose reference winding number is outside the |w - 0.5| <= band zone (BASELINE.json north_star)."""
    return np.abs(np.asarray(w_ref, dtype=np.float64) - 0.5) > band
