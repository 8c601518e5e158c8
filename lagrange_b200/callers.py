"""The reference's two example callers of the engine, as batched operations.

    fix_orientation        modules/winding/examples/fix_orientation.cpp:86-115
    sample_points_in_mesh  modules/winding/examples/sample_points_in_mesh.cpp:58-76

The reference issues one `solid_angle` / `is_inside` call per facet / per sample; here the query points of the whole mesh
are generated with the same float arithmetic and sent through the engine in one batch. Everything that decides a result
(the points, the criterion, the threshold comparison, the random stream) follows the reference line by line, so that with
equal solid angles the outputs are identical.
"""
from __future__ import annotations

import numpy as np

from .winding import FastWindingNumber

_PI = 3.14159265358979323846  # lagrange::internal::pi is a double (core/include/lagrange/internal/constants.h:16)


def orientation_probe_points(vertices, facets, epsilon=1e-2):
    """The two query points of every facet: barycenter +/- epsilon * unit normal, float32 like the Eigen::Vector3f code
    (fix_orientation.cpp:88-99). Returns (pp [F,3], qq [F,3])."""
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    f = np.asarray(facets).reshape(-1, 3)
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    n = np.cross(b - a, c - a).astype(np.float32)
    norm = np.sqrt((n * n).sum(axis=1, dtype=np.float32), dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        n = (n / norm[:, None]).astype(np.float32)  # a degenerate facet gives NaN, as Eigen's normalized() does
    bary = ((a + b + c) / np.float32(3)).astype(np.float32)
    eps = np.float32(epsilon)
    return (bary + eps * n).astype(np.float32), (bary - eps * n).astype(np.float32)


def orientation_criterion(omega_pp, omega_qq):
    """(solid_angle(pp) - solid_angle(qq)) / (4.f * pi): float difference, double division (fix_orientation.cpp:100-101)."""
    d = np.asarray(omega_pp, dtype=np.float32) - np.asarray(omega_qq, dtype=np.float32)
    return d.astype(np.float64) / (4.0 * _PI)


def fix_orientation(vertices, facets, engine: FastWindingNumber, epsilon=1e-2, threshold=0.8):
    """Flip (swap corners 0 and 1 of) every facet of the input mesh whose orientation disagrees with the reference mesh the
    engine was built from. Returns (facets_out, criterion [F] float64, counts dict) — fix_orientation.cpp:86-115."""
    f = np.array(facets, copy=True).reshape(-1, 3)
    pp, qq = orientation_probe_points(vertices, f, epsilon)
    omega = engine.solid_angle(np.concatenate([pp, qq], axis=0))
    nf = len(f)
    crit = orientation_criterion(omega[:nf], omega[nf:])
    thr = float(np.float32(threshold))
    flip = crit > thr  # NaN (degenerate facet) compares false, as in the reference
    f[flip, 0], f[flip, 1] = f[flip, 1].copy(), f[flip, 0].copy()
    counts = {"positive": int(flip.sum()), "negative": int((crit < thr).sum()), "total": nf}
    return f, crit, counts


def mt19937_uniform_float(n, lo, hi, state=None):
    """`std::uniform_real_distribution<float>(lo, hi)(gen)` for a default-seeded `std::mt19937` as libstdc++ computes it:
    one 32-bit draw per value, canonical = float(draw) / 2^32 (clamped below 1), value = canonical * (hi - lo) + lo, all in
    float. `state` is a numpy RandomState positioned in the stream (default: seed 5489 = std::mt19937's default seed).
    Checked against g++'s own output in tests/golden/mt19937_uniform_float.npz."""
    rs = state if state is not None else np.random.RandomState(5489)
    raw = rs._bit_generator.random_raw(n).astype(np.uint32)
    canon = raw.astype(np.float32) / np.float32(4294967296.0)
    canon = np.minimum(canon, np.nextafter(np.float32(1.0), np.float32(0.0)))
    lo32, hi32 = np.float32(lo), np.float32(hi)
    return (canon * (hi32 - lo32) + lo32).astype(np.float32)


def sample_points_in_mesh(engine: FastWindingNumber, bbox_min, bbox_max, num_samples=10000):
    """Rejection-sample the mesh interior: uniform points in the bounding box from a default-seeded mt19937 (x, y, z drawn in
    that order per point), kept when `is_inside` (sample_points_in_mesh.cpp:58-76). Returns the kept points [K,3] float32 in
    draw order."""
    rs = np.random.RandomState(5489)
    raw = mt19937_uniform_float(3 * int(num_samples), 0.0, 1.0, rs).reshape(-1, 3)  # canonical draws, consumed x, y, z
    lo = np.asarray(bbox_min, dtype=np.float32)
    hi = np.asarray(bbox_max, dtype=np.float32)
    pts = (raw * (hi - lo)[None, :] + lo[None, :]).astype(np.float32)
    keep = engine.is_inside(pts).astype(bool)
    return pts[keep]


# ---- consumers of the closest-point query -------------------------------------------------
#     compute_mesh_distances / compute_hausdorff / compute_chamfer   modules/bvh/src/compute_mesh_distances.cpp:45-166
# The reference builds a TriangleAABBTree over the target and calls get_closest_point once per source vertex from a TBB
# parallel_for (:60-73); here the source's vertices go through wn_closest_point in one batch on the target's engine.
def _target_engine(target_vertices, target_facets, engine):
    return engine if engine is not None else FastWindingNumber(target_vertices, target_facets)


def compute_mesh_distances(source_vertices, target_vertices, target_facets, engine=None) -> np.ndarray:
    """Distance from every source vertex to the closest point of the target mesh (float32 [nV]); an empty target gives zeros,
    like the reference (compute_mesh_distances.cpp:53-56)."""
    sv = np.ascontiguousarray(source_vertices, dtype=np.float32).reshape(-1, 3)
    if len(np.asarray(target_facets).reshape(-1, 3)) == 0 or len(sv) == 0:
        return np.zeros(len(sv), dtype=np.float32)
    sq, _, _ = _target_engine(target_vertices, target_facets, engine).closest_point(sv)
    return np.sqrt(sq)


def compute_hausdorff(source_vertices, source_facets, target_vertices, target_facets) -> float:
    """max(directed source->target, directed target->source) over the vertices (compute_mesh_distances.cpp:100-126)."""
    fwd = compute_mesh_distances(source_vertices, target_vertices, target_facets)
    bwd = compute_mesh_distances(target_vertices, source_vertices, source_facets)
    return float(max(fwd.max() if len(fwd) else 0.0, bwd.max() if len(bwd) else 0.0))


def compute_chamfer(source_vertices, source_facets, target_vertices, target_facets) -> float:
    """mean squared source->target distance + mean squared target->source distance (compute_mesh_distances.cpp:128-166)."""
    fwd = compute_mesh_distances(source_vertices, target_vertices, target_facets).astype(np.float64)
    bwd = compute_mesh_distances(target_vertices, source_vertices, source_facets).astype(np.float64)
    out = 0.0
    if len(fwd):
        out += float((fwd * fwd).sum() / len(fwd))
    if len(bwd):
        out += float((bwd * bwd).sum() / len(bwd))
    return out
