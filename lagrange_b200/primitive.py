"""Synthetic input meshes and query sets for the FastWindingNumber hot path.

These re-specify, in numpy, just enough of ``lagrange::primitive`` to produce the benchmark inputs; they are input
generators, not a port of the primitive module (out of scope as code):

* torus               : modules/primitive/src/generate_torus.cpp:42-148 (ring in the XZ plane, R=5, r=1 defaults,
                        ``triangulate`` => centroid fan, core/src/triangulate_polygonal_facets.cpp:302-360)
* icosahedron         : modules/primitive/src/generate_icosahedron.cpp:31-70 (unit circumsphere)
* subdivided sphere   : modules/primitive/src/generate_subdivided_sphere.cpp:27-84 (subdivide, normalise, scale);
                        OpenSubdiv's Loop scheme is replaced by midpoint subdivision (every vertex is re-projected
                        onto the sphere anyway)
* cell-centred lattice: modules/volume/src/mesh_to_volume.cpp:147-149  p = voxel_size * (ijk + 1/2)

All meshes are returned as ``(vertices float32 [nV,3], facets int32 [nF,3])`` with outward orientation, which is
what ``FastWindingNumber`` converts its input to (modules/winding/src/FastWindingNumber.cpp:40-52).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "generate_torus", "generate_icosahedron", "generate_octahedron", "generate_subdivided_sphere", "midpoint_subdivide",
    "make_soup", "lattice_for_bbox", "lattice_points", "near_surface_points", "uniform_points_in_bbox", "mesh_bbox",
    "config_mesh", "config_queries",
]


def mesh_bbox(vertices):
    v = np.asarray(vertices)
    return v.min(axis=0), v.max(axis=0)


def generate_torus(major_radius=5.0, minor_radius=1.0, ring_segments=100, pipe_segments=50, triangulate=True):
    """Torus swept around the Y axis. ``triangulate`` splits every quad into a 4-triangle centroid fan.

    ring_segments x pipe_segments quads -> 4x triangles, nV = 2 * ring * pipe (quad corners + quad centroids).
    """
    R, r, nr, npipe = float(major_radius), float(minor_radius), int(ring_segments), int(pipe_segments)
    u = (np.arange(nr, dtype=np.float64) / nr) * 2 * np.pi  # around the ring
    w = (np.arange(npipe, dtype=np.float64) / npipe) * 2 * np.pi  # around the pipe
    uu, ww = np.meshgrid(u, w, indexing="ij")
    rad = R + r * np.cos(ww)
    P = np.stack([rad * np.cos(uu), r * np.sin(ww), rad * np.sin(uu)], axis=-1).reshape(-1, 3)

    i = np.arange(nr)[:, None]
    j = np.arange(npipe)[None, :]
    v00 = (i * npipe + j).ravel()
    v10 = (((i + 1) % nr) * npipe + j).ravel()
    v11 = (((i + 1) % nr) * npipe + (j + 1) % npipe).ravel()
    v01 = (i * npipe + (j + 1) % npipe).ravel()
    # quad loop chosen so that the normal points out of the tube (checked in tests: w(ring centre line) == 1)
    quad = np.stack([v00, v01, v11, v10], axis=1)
    if not triangulate:
        # two triangles per quad
        F = np.concatenate([quad[:, [0, 1, 2]], quad[:, [0, 2, 3]]], axis=0)
        return P.astype(np.float32), F.astype(np.int32)
    cid = len(P) + np.arange(len(quad))
    C = P[quad].mean(axis=1)
    V = np.concatenate([P, C], axis=0)
    fans = [np.stack([cid, quad[:, k], quad[:, (k + 1) % 4]], axis=1) for k in range(4)]
    F = np.stack(fans, axis=1).reshape(-1, 3)
    return V.astype(np.float32), F.astype(np.int32)


def generate_icosahedron(radius=1.0):
    """Regular icosahedron on a sphere of ``radius`` (golden-ratio construction, outward oriented)."""
    t = (1.0 + np.sqrt(5.0)) / 2.0
    V = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    F = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
                  [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    F = _orient_outward(V, F)
    return (V * radius).astype(np.float32), F.astype(np.int32)


def generate_octahedron(radius=1.0):
    V = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64)
    F = np.array([[0, 2, 4], [2, 1, 4], [1, 3, 4], [3, 0, 4], [2, 0, 5], [1, 2, 5], [3, 1, 5], [0, 3, 5]], dtype=np.int64)
    F = _orient_outward(V, F)
    return (V * radius).astype(np.float32), F.astype(np.int32)


def _orient_outward(V, F):
    """Flip triangles of a star-shaped (about the origin) mesh whose normal points inward."""
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    n = np.cross(b - a, c - a)
    flip = np.einsum("ij,ij->i", n, (a + b + c)) < 0
    F = F.copy()
    F[flip] = F[flip][:, [0, 2, 1]]
    return F


def midpoint_subdivide(V, F):
    """One 1-to-4 midpoint split. Keeps orientation. V float64 in, float64 out."""
    F = np.asarray(F, dtype=np.int64)
    nV = len(V)
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], axis=0)
    e.sort(axis=1)
    key = e[:, 0] * nV + e[:, 1]
    uniq, inv = np.unique(key, return_inverse=True)
    mid = 0.5 * (V[uniq // nV] + V[uniq % nV])
    nF = len(F)
    m01, m12, m20 = nV + inv[:nF], nV + inv[nF:2 * nF], nV + inv[2 * nF:]
    V2 = np.concatenate([V, mid], axis=0)
    F2 = np.concatenate([
        np.stack([F[:, 0], m01, m20], axis=1),
        np.stack([F[:, 1], m12, m01], axis=1),
        np.stack([F[:, 2], m20, m12], axis=1),
        np.stack([m01, m12, m20], axis=1),
    ], axis=0)
    return V2, F2


def generate_subdivided_sphere(base="icosahedron", subdiv_level=0, radius=1.0):
    """Subdivide ``base`` (``'icosahedron'`` | ``'octahedron'``) ``subdiv_level`` times, re-projecting on the sphere.

    icosahedron level 8 -> 1 310 720 triangles / 655 362 vertices (cfg2);
    octahedron  level 10 -> 8 388 608 triangles (cfg4).
    """
    V, F = (generate_icosahedron if base == "icosahedron" else generate_octahedron)(1.0)
    V, F = V.astype(np.float64), F.astype(np.int64)
    for _ in range(int(subdiv_level)):
        V, F = midpoint_subdivide(V, F)
        V /= np.linalg.norm(V, axis=1, keepdims=True)
    return (V * radius).astype(np.float32), F.astype(np.int32)


def make_soup(V, F, seed=0xC0FFEE03, n_caps=24, cap_radius=0.15, dup_frac=0.01, flip_frac=0.01):
    """cfg3: punch holes (delete triangles whose centroid direction lies in random angular caps), duplicate 1 % and
    flip 1 % of the triangles -> open, non-manifold, inconsistently oriented soup."""
    rng = np.random.Generator(np.random.PCG64(seed))
    V = np.asarray(V, dtype=np.float32)
    F = np.asarray(F, dtype=np.int32)
    c = V[F].mean(axis=1).astype(np.float64)
    c -= c.mean(axis=0)
    d = c / np.maximum(np.linalg.norm(c, axis=1, keepdims=True), 1e-30)
    caps = rng.normal(size=(n_caps, 3))
    caps /= np.linalg.norm(caps, axis=1, keepdims=True)
    keep = np.ones(len(F), dtype=bool)
    cos_r = np.cos(cap_radius)
    for k in range(n_caps):
        keep &= (d @ caps[k]) < cos_r
    F = F[keep]
    n = len(F)
    dup = rng.choice(n, size=int(round(n * dup_frac)), replace=False)
    F = np.concatenate([F, F[dup]], axis=0)
    flip = rng.choice(len(F), size=int(round(n * flip_frac)), replace=False)
    F[flip] = F[flip][:, [0, 2, 1]]
    return V, np.ascontiguousarray(F)


def lattice_for_bbox(lo, hi, n, inflate=0.05):
    """Cell-centred n^3 (or (nx,ny,nz)) lattice covering the bbox inflated by ``inflate`` of its extent per side.

    Returns (origin[3], spacing[3], dims[3]); point (i,j,k) = origin + spacing * (ijk + 0.5), x fastest.
    """
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    ext = hi - lo
    lo, hi = lo - inflate * ext, hi + inflate * ext
    dims = np.array([n, n, n] if np.isscalar(n) else n, dtype=np.int64)
    spacing = (hi - lo) / dims
    return lo.astype(np.float32), spacing.astype(np.float32), dims


def lattice_points(origin, spacing, dims, first=0, count=None, stride=1):
    """Materialise lattice points in float32 with exactly the arithmetic the kernels use:
    p = origin + spacing * (float(i) + 0.5f), each op rounded to float32."""
    dims = np.asarray(dims, dtype=np.int64)
    total = int(dims[0] * dims[1] * dims[2])
    if count is None:
        count = (total - first + stride - 1) // stride
    idx = first + np.arange(count, dtype=np.int64) * stride
    ix, iy, iz = idx % dims[0], (idx // dims[0]) % dims[1], idx // (dims[0] * dims[1])
    o, s = np.asarray(origin, dtype=np.float32), np.asarray(spacing, dtype=np.float32)
    half = np.float32(0.5)
    out = np.empty((count, 3), dtype=np.float32)
    for a, ii in enumerate((ix, iy, iz)):
        out[:, a] = o[a] + s[a] * (ii.astype(np.float32) + half)
    return out


def near_surface_points(V, F, n, sigma_rel=1e-3, seed=0xC0FFEE04):
    """cfg4: area-uniform surface samples displaced along the normal by N(0, (sigma_rel * bbox diagonal)^2), random order."""
    rng = np.random.Generator(np.random.PCG64(seed))
    V = np.asarray(V, dtype=np.float64)
    F = np.asarray(F)
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    nrm = np.cross(b - a, c - a)
    area = 0.5 * np.linalg.norm(nrm, axis=1)
    cdf = np.cumsum(area)
    t = np.searchsorted(cdf, rng.random(n) * cdf[-1], side="right").clip(0, len(F) - 1)
    r1, r2 = np.sqrt(rng.random(n)), rng.random(n)
    w0, w1, w2 = 1 - r1, r1 * (1 - r2), r1 * r2
    p = w0[:, None] * a[t] + w1[:, None] * b[t] + w2[:, None] * c[t]
    nh = nrm[t] / np.maximum(np.linalg.norm(nrm[t], axis=1, keepdims=True), 1e-300)
    lo, hi = mesh_bbox(V)
    sigma = sigma_rel * np.linalg.norm(hi - lo)
    p += nh * rng.normal(0.0, sigma, size=(n, 1))
    return p.astype(np.float32)


def uniform_points_in_bbox(lo, hi, n, inflate=0.05, seed=0xC0FFEE05):
    rng = np.random.Generator(np.random.PCG64(seed))
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    ext = hi - lo
    lo, hi = lo - inflate * ext, hi + inflate * ext
    return (lo + rng.random((n, 3)) * (hi - lo)).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# The five BASELINE.json configs. ``scale`` < 1 shrinks them for parity tests.
# ---------------------------------------------------------------------------------------------------------------
def config_mesh(cfg: int, small: bool = False):
    if cfg == 1:
        return generate_torus(5.0, 1.0, 100, 50) if not small else generate_torus(5.0, 1.0, 24, 12)
    if cfg == 2:
        return generate_subdivided_sphere("icosahedron", 8 if not small else 3)
    if cfg == 3:
        V, F = generate_torus(5.0, 1.0, 250, 200) if not small else generate_torus(5.0, 1.0, 40, 24)
        return make_soup(V, F, seed=0xC0FFEE03)
    if cfg == 4:
        return generate_subdivided_sphere("octahedron", 10 if not small else 4)
    if cfg == 5:
        return generate_torus(5.0, 1.0, 250, 100) if not small else generate_torus(5.0, 1.0, 30, 16)
    raise ValueError(f"unknown config {cfg}")


def config_queries(cfg: int, V, F, small: bool = False):
    """Returns ('grid', (origin, spacing, dims)) or ('points', float32 [n,3])."""
    lo, hi = mesh_bbox(V)
    if cfg == 1:
        return "grid", lattice_for_bbox(lo, hi, 100 if not small else 24)
    if cfg == 2:
        n = 512 if not small else 32
        o = np.full(3, -1.1, dtype=np.float32)
        return "grid", (o, np.full(3, 2.2 / n, dtype=np.float32), np.array([n, n, n], dtype=np.int64))
    if cfg == 3:
        return "grid", lattice_for_bbox(lo, hi, 256 if not small else 28)
    if cfg == 4:
        return "points", near_surface_points(V, F, (64 << 20) if not small else 20000, seed=0xC0FFEE04)
    if cfg == 5:
        return "points", uniform_points_in_bbox(lo, hi, (1 << 24) if not small else 20000, seed=0xC0FFEE05)
    raise ValueError(f"unknown config {cfg}")
