"""Minimal Wavefront OBJ reader/writer: enough to feed the engine the meshes the reference's
tests and examples load with `lagrange::io::load_mesh` (modules/winding/tests/test_fast_winding_number.cpp:90,
modules/winding/examples/*.cpp). Positions and faces only; `v/vt/vn` corner syntax and negative (relative) indices are
understood; polygons are kept, or fan-triangulated on request (the engine itself only accepts triangles, like the reference:
FastWindingNumber.cpp:91-96)."""
from __future__ import annotations

import numpy as np

from .mesh import SurfaceMesh
from .winding import Error


def load_obj(path, triangulate=False, scalar=np.float32, index=np.uint32) -> SurfaceMesh:
    verts = []
    faces = []
    with open(path, "r", encoding="utf-8", errors="replace") as fh:
        for lineno, line in enumerate(fh, 1):
            if not line or line[0] not in "vf":
                continue
            parts = line.split()
            if not parts:
                continue
            if parts[0] == "v":
                if len(parts) < 4:
                    raise Error(f"{path}:{lineno}: vertex needs three coordinates")
                verts.append((float(parts[1]), float(parts[2]), float(parts[3])))
            elif parts[0] == "f":
                if len(parts) < 4:
                    raise Error(f"{path}:{lineno}: face needs at least three corners")
                corner = []
                for tok in parts[1:]:
                    i = int(tok.split("/", 1)[0])
                    i = i - 1 if i > 0 else len(verts) + i  # negative indices count back from the last vertex read so far
                    if i < 0 or i >= len(verts):
                        raise Error(f"{path}:{lineno}: vertex index out of range")
                    corner.append(i)
                faces.append(corner)
    mesh = SurfaceMesh(3, scalar, index)
    if verts:
        mesh.add_vertices(np.asarray(verts, dtype=np.float64))
    for c in faces:
        if len(c) == 3:
            mesh.add_triangle(*c)
        elif triangulate:
            for k in range(1, len(c) - 1):
                mesh.add_triangle(c[0], c[k], c[k + 1])
        else:
            mesh.add_polygon(c)
    return mesh


def save_obj(path, vertices, facets):
    v = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
    with open(path, "w", encoding="utf-8") as fh:
        for p in v:
            fh.write(f"v {p[0]:.9g} {p[1]:.9g} {p[2]:.9g}\n")
        for f in facets:
            fh.write("f " + " ".join(str(int(i) + 1) for i in f) + "\n")
