"""Minimal stand-in for ``lagrange::SurfaceMesh<Scalar, Index>`` — only what FastWindingNumber's constructor touches
(modules/core/include/lagrange/SurfaceMesh.h: add_vertices :259, add_triangles :300, get_dimension :1918,
is_triangle_mesh :632, get_num_vertices/facets; contiguous row-major V/F buffers as vertex_view/facet_view expose them,
modules/core/src/views.cpp:156-175). The full mesh library is out of scope."""
from __future__ import annotations

import numpy as np


class SurfaceMesh:
    def __init__(self, dimension: int = 3, scalar=np.float32, index=np.uint32):
        self._dim = int(dimension)
        self._scalar = np.dtype(scalar)
        self._index = np.dtype(index)
        self._v = np.zeros((0, self._dim), dtype=self._scalar)
        self._facets = []  # list of index arrays (polygons of any size)

    def add_vertex(self, p):
        self.add_vertices(np.asarray(p).reshape(1, -1))

    def add_vertices(self, points):
        p = np.asarray(points, dtype=self._scalar).reshape(-1, self._dim)
        self._v = np.concatenate([self._v, p], axis=0)

    def add_triangle(self, a, b, c):
        self._facets.append(np.array([a, b, c], dtype=self._index))

    def add_triangles(self, tris):
        for t in np.asarray(tris, dtype=self._index).reshape(-1, 3):
            self._facets.append(t)

    def add_polygon(self, idx):
        self._facets.append(np.asarray(idx, dtype=self._index))

    def get_dimension(self) -> int:
        return self._dim

    def get_num_vertices(self) -> int:
        return len(self._v)

    def get_num_facets(self) -> int:
        return len(self._facets)

    def is_triangle_mesh(self) -> bool:
        # an empty mesh is regular with no corners and counts as a triangle mesh (core/src/SurfaceMesh.cpp:2320-2323)
        return all(len(f) == 3 for f in self._facets)

    @property
    def vertices(self) -> np.ndarray:
        return self._v

    @property
    def facets(self) -> np.ndarray:
        if not self._facets:
            return np.zeros((0, 3), dtype=self._index)
        return np.stack(self._facets, axis=0)

    @classmethod
    def from_arrays(cls, vertices, facets, scalar=None, index=None):
        v = np.asarray(vertices)
        f = np.asarray(facets)
        m = cls(v.shape[1], scalar or v.dtype, index or f.dtype)
        m._v = np.ascontiguousarray(v, dtype=m._scalar)
        m._facets = list(np.ascontiguousarray(f, dtype=m._index))
        return m
