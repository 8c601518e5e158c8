"""Narrow-band signed distance (K10, wn_sdf_grid) and the mesh_to_volume mirror.
Ground truth: oracle.distance64 (double, brute force, a different point-triangle formulation than the kernel's)."""
import numpy as np
import pytest


def test_oracle_distance_known_answers(oracle_mod):
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32)
    F = np.array([[0, 1, 2]], dtype=np.int32)
    q = np.array([[0.25, 0.25, 2.0],   # above the interior: plane distance
                  [-3.0, -4.0, 0.0],   # nearest feature: vertex a
                  [0.5, -2.0, 0.0],    # edge ab
                  [2.0, 2.0, 0.0],     # edge bc: distance to the line x + y = 1
                  [0.2, 0.2, 0.0]], dtype=np.float32)
    d = oracle_mod.distance64(V, F, q)
    assert np.allclose(d, [2.0, 5.0, 2.0, 3.0 / np.sqrt(2.0), 0.0], atol=1e-12)


def test_device_point_triangle_routine_matches_oracle(emul_mod, oracle_mod):
    rng = np.random.RandomState(5)
    n = 20000
    tris = rng.randn(n, 3, 3).astype(np.float32)
    tris[:200, 2] = tris[:200, 1]                       # degenerate: two equal corners
    tris[200:400, 2] = 0.5 * (tris[200:400, 0] + tris[200:400, 1])  # degenerate: collinear
    pts = (rng.randn(n, 3) * 2).astype(np.float32)
    pts[400:600] = tris[400:600, 0]                     # on a vertex
    got = np.sqrt(emul_mod.point_tri_dist2(pts, tris).astype(np.float64))
    want = np.array([oracle_mod.distance64(tris[i], np.array([[0, 1, 2]], dtype=np.int32), pts[i:i + 1])[0] for i in range(0, n, 7)])
    assert np.abs(got[::7] - want).max() < 2e-5 * max(1.0, want.max())


def test_voxel_size_and_index_box(prim):
    from lagrange_b200 import volume
    from lagrange_b200.winding import Error

    V, F = prim.generate_torus(5, 1, 16, 8)
    diag = np.linalg.norm(V.max(axis=0).astype(np.float64) - V.min(axis=0))
    assert volume.resolve_voxel_size(V, -0.01) == pytest.approx(0.01 * diag)
    assert volume.resolve_voxel_size(V, 0.25) == 0.25
    with pytest.raises(Error):
        volume.resolve_voxel_size(V, 0.0)
    lo, dims = volume.band_index_box(V, 0.25)
    centres_lo = 0.25 * (lo + 0.5)
    centres_hi = 0.25 * (lo + dims - 1 + 0.5)
    band = 3 * 0.25
    assert np.all(centres_lo <= V.min(axis=0) - band) and np.all(centres_hi >= V.max(axis=0) + band)


@pytest.mark.gpu
@pytest.mark.parametrize("leaf_size", [1, 8])
def test_sdf_grid_matches_brute_force(prim, oracle_mod, leaf_size):
    import lagrange_b200 as lb

    V, F = prim.generate_torus(5, 1, 40, 20)
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh", leaf_size=leaf_size)
    origin, spacing, dims = prim.lattice_for_bbox(V.min(axis=0), V.max(axis=0), (37, 19, 41), inflate=0.1)
    band = 4.5 * float(spacing[0])
    sdf, active = eng.sdf_grid(origin, spacing, dims, band)
    q = prim.lattice_points(origin, spacing, dims)
    d = oracle_mod.distance64(V, F, q).reshape(sdf.shape)
    w = (oracle_mod.exact64(V, F, q) / (4 * np.pi)).reshape(sdf.shape)
    assert np.abs(np.abs(sdf) - np.minimum(d, band)).max() < 1e-5
    clear = np.abs(w - 0.5) > 5e-3
    assert np.array_equal((sdf < 0)[clear], (w > 0.5)[clear])
    assert active == int((np.abs(sdf) < np.float32(band)).sum())
    assert abs(active - int((d < band).sum())) <= 4
    # unsigned variant
    usdf, uactive = eng.sdf_grid(origin, spacing, dims, band, signed=False)
    assert np.array_equal(usdf, np.abs(sdf)) and uactive == active


@pytest.mark.gpu
def test_sdf_grid_open_soup_and_device_output(prim, oracle_mod):
    import torch

    import lagrange_b200 as lb

    V, F = prim.config_mesh(3, small=True)  # open, non-manifold soup
    eng = lb.FastWindingNumber(V, F, hierarchy="lbvh")
    origin, spacing, dims = prim.lattice_for_bbox(V.min(axis=0), V.max(axis=0), (33, 17, 29), inflate=0.05)
    band = 3.0 * float(max(spacing))
    out = torch.empty(int(np.prod(dims)), dtype=torch.float32, device="cuda")
    sdf, active = eng.sdf_grid(origin, spacing, dims, band, out=out)
    sdf = sdf.cpu().numpy()
    q = prim.lattice_points(origin, spacing, dims)
    d = oracle_mod.distance64(V, F, q).reshape(sdf.shape)
    assert np.abs(np.abs(sdf) - np.minimum(d, band)).max() < 1e-5
    inside = eng.is_inside_grid(origin, spacing, dims).reshape(sdf.shape)
    assert np.array_equal(sdf < 0, inside.astype(bool))


@pytest.mark.gpu
def test_mesh_to_volume_mirror(prim, oracle_mod):
    import lagrange_b200 as lb
    from lagrange_b200 import volume

    V, F = prim.generate_subdivided_sphere("icosahedron", 3)
    grid = volume.mesh_to_volume((V, F), volume.MeshToVolumeOptions(voxel_size=-0.02, signing_method="WindingNumber"))
    vs = grid.voxel_size
    assert vs == pytest.approx(0.02 * 2 * np.sqrt(3.0), rel=1e-3)
    assert grid.background == pytest.approx(3 * vs, rel=1e-6)
    nz, ny, nx = grid.values.shape
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    centres = grid.index_to_world(np.stack([i, j, k], axis=-1) + np.asarray(grid.ijk_min))
    r = np.linalg.norm(centres, axis=-1)
    # a unit sphere (facetted, level 3: sagitta ~ 1e-2): the level set is r - 1 inside the band, +-background outside
    band = np.abs(grid.values) < grid.background * 0.999
    assert np.abs(grid.values[band] - (r[band] - 1.0)).max() < 2.5e-2
    assert np.all(grid.values[r < 1 - 3.2 * vs] == -np.float32(grid.background))
    assert np.all(grid.values[r > 1 + 3.2 * vs] == np.float32(grid.background))
    assert grid.active_voxels == int(band.sum()) or abs(grid.active_voxels - int((np.abs(grid.values) < grid.background).sum())) == 0
    un = volume.mesh_to_volume((V, F), volume.MeshToVolumeOptions(voxel_size=vs, signing_method="Unsigned"))
    assert np.array_equal(un.values, np.abs(grid.values))
    with pytest.raises(lb.Error):
        volume.mesh_to_volume((V, F), volume.MeshToVolumeOptions(signing_method="FloodFill"))
