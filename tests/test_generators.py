"""CPU tier: the synthetic inputs of BASELINE.md section 4."""
import numpy as np

from conftest import FOUR_PI


def _edges_closed_and_oriented(F):
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
    key = e[:, 0].astype(np.int64) * (F.max() + 1) + e[:, 1]
    rev = e[:, 1].astype(np.int64) * (F.max() + 1) + e[:, 0]
    return len(np.unique(key)) == len(key) and np.array_equal(np.sort(key), np.sort(rev))


def _volume(V, F):
    a, b, c = V[F[:, 0]].astype(np.float64), V[F[:, 1]].astype(np.float64), V[F[:, 2]].astype(np.float64)
    return np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0


def test_cfg1_torus_counts_and_orientation(prim):
    V, F = prim.config_mesh(1)
    assert V.shape == (10000, 3) and F.shape == (20000, 3) and V.dtype == np.float32 and F.dtype == np.int32
    assert _edges_closed_and_oriented(F)
    assert abs(_volume(V, F) - 2 * np.pi**2 * 5.0) / (2 * np.pi**2 * 5.0) < 5e-3  # outward: positive volume 2 pi^2 R r^2
    kind, (o, s, d) = prim.config_queries(1, V, F)
    assert kind == "grid" and tuple(d) == (100, 100, 100)
    lo, hi = prim.mesh_bbox(V)
    assert np.allclose(o, lo - 0.05 * (hi - lo), atol=1e-5) and np.allclose(o + s * d, hi + 0.05 * (hi - lo), atol=1e-4)


def test_subdivided_spheres(prim):
    V, F = prim.generate_subdivided_sphere("icosahedron", 4)
    assert F.shape == (20 * 4**4, 3) and V.shape == (10 * 4**4 + 2, 3)  # level 8 -> 1 310 720 / 655 362 (cfg2)
    assert np.allclose(np.linalg.norm(V, axis=1), 1, atol=1e-6) and _edges_closed_and_oriented(F) and _volume(V, F) > 4.0
    V, F = prim.generate_subdivided_sphere("octahedron", 3)
    assert F.shape == (8 * 4**3, 3)  # level 10 -> 8 388 608 (cfg4)
    assert _edges_closed_and_oriented(F) and _volume(V, F) > 4.0


def test_soup_is_open_duplicated_and_flipped(prim):
    V, F0 = prim.generate_torus(5, 1, 60, 40)
    _, F = prim.make_soup(V, F0, seed=0xC0FFEE03)
    assert len(F) < len(F0) and not _edges_closed_and_oriented(F)
    _, F2 = prim.make_soup(V, F0, seed=0xC0FFEE03)
    assert np.array_equal(F, F2)  # seeded
    s = np.sort(np.sort(F, axis=1), axis=0)
    assert len(np.unique(np.sort(F, axis=1), axis=0)) < len(F)  # duplicates exist


def test_lattice_points_are_cell_centred(prim, emul_mod):
    o, s, d = prim.lattice_for_bbox([-1, -2, -3], [1, 2, 3], (4, 5, 6), inflate=0.0)
    P = prim.lattice_points(o, s, d)
    assert P.shape == (120, 3) and np.allclose(P[0], o + 0.5 * s) and np.allclose(P[-1], o + s * (d - 0.5))
    assert np.allclose(P[1] - P[0], [s[0], 0, 0])  # x fastest
    # identical to the device's lattice arithmetic (wn_lattice_coord), bit for bit
    L = emul_mod.lib()
    for i in (0, 1, 3):
        assert P[i, 0] == np.float32(L.emul_lattice_coord(float(o[0]), float(s[0]), i))
    assert np.array_equal(prim.lattice_points(o, s, d, first=7, stride=5), P[7::5])


def test_near_surface_and_uniform_points_are_seeded(prim):
    V, F = prim.generate_torus(5, 1, 30, 16)
    a = prim.near_surface_points(V, F, 1000, seed=1)
    assert np.array_equal(a, prim.near_surface_points(V, F, 1000, seed=1)) and a.dtype == np.float32
    d = np.abs(np.sqrt((np.sqrt(a[:, 0] ** 2 + a[:, 2] ** 2) - 5) ** 2 + a[:, 1] ** 2) - 1)
    assert d.mean() < 0.05  # hugging the surface
    u = prim.uniform_points_in_bbox(*prim.mesh_bbox(V), 1000, seed=2)
    assert u.shape == (1000, 3)
