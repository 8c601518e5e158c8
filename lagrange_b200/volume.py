"""mesh_to_volume with winding-number signing on the GPU.

Mirrors `lagrange::volume::mesh_to_volume(mesh, MeshToVolumeOptions)` for `Sign::WindingNumber` and `Sign::Unsigned`
(modules/volume/include/lagrange/volume/mesh_to_volume.h:26-60, modules/volume/src/mesh_to_volume.cpp:129-183): the same
voxel-size rule, the same cell-centred transform, 3 voxels of band on either side, the interior test evaluated at every
voxel centre. The reference returns an OpenVDB level set (sparse); OpenVDB does not exist here, so the result is the dense
block of voxels that can hold band values, plus what is needed to place it: `VolumeGrid`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .mesh import SurfaceMesh
from .winding import Error, FastWindingNumber

EXTERIOR_BANDWIDTH = 3.0  # voxels, mesh_to_volume.cpp:161-162
INTERIOR_BANDWIDTH = 3.0


@dataclass
class MeshToVolumeOptions:
    """volume/mesh_to_volume.h:26-44. signing_method: "WindingNumber" or "Unsigned" ("FloodFill" is OpenVDB's own method
    and is not provided)."""

    voxel_size: float = -0.01  # negative: relative to the bbox diagonal
    signing_method: str = "WindingNumber"


@dataclass
class VolumeGrid:
    """Dense narrow-band level set. values[k, j, i] belongs to voxel ijk = ijk_min + (i, j, k), whose centre is
    voxel_size * (ijk + 1/2) (the reference's linear transform post-translated by half a voxel). |values| == background
    outside the band; negative inside."""

    values: object  # numpy [nz, ny, nx] float32, or a torch CUDA tensor of that shape
    ijk_min: tuple
    voxel_size: float
    background: float
    active_voxels: int

    def index_to_world(self, ijk):
        return self.voxel_size * (np.asarray(ijk, dtype=np.float64) + 0.5)


def resolve_voxel_size(vertices, voxel_size):
    """mesh_to_volume.cpp:131-145: a negative size is relative to the bbox diagonal."""
    if voxel_size < 0:
        v = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
        diag = float(np.linalg.norm(v.max(axis=0) - v.min(axis=0))) if len(v) else 0.0
        voxel_size = abs(voxel_size) * diag
    if not (voxel_size > 0) or not math.isfinite(voxel_size):
        raise Error(f"Voxel size too small: {voxel_size}")
    return float(voxel_size)


def band_index_box(vertices, voxel_size, band_voxels=EXTERIOR_BANDWIDTH):
    """Smallest index box whose voxel centres include every point within the band of the mesh's bounding box."""
    v = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
    pad = int(math.ceil(band_voxels)) + 1
    lo = np.floor(v.min(axis=0) / voxel_size).astype(np.int64) - pad
    hi = np.floor(v.max(axis=0) / voxel_size).astype(np.int64) + pad
    return lo, hi - lo + 1


def mesh_to_volume(mesh, options: MeshToVolumeOptions | None = None, engine: FastWindingNumber | None = None, device_output=False):
    """Returns a VolumeGrid. `mesh`: SurfaceMesh or (vertices, facets). An existing engine for the same mesh may be passed."""
    options = options or MeshToVolumeOptions()
    if options.signing_method not in ("WindingNumber", "Unsigned"):
        raise Error(f"signing method {options.signing_method!r} is not provided (WindingNumber, Unsigned)")
    if not isinstance(mesh, SurfaceMesh):
        mesh = SurfaceMesh.from_arrays(np.asarray(mesh[0]), np.asarray(mesh[1]))
    if not mesh.is_triangle_mesh():
        # the reference triangulates polygonal facets first (mesh_to_volume.cpp:100-106); the mesh library is out of scope here
        raise Error("mesh_to_volume: triangulate the mesh first (only triangle meshes are supported)")
    V = mesh.vertices
    if len(V) == 0 or mesh.get_num_facets() == 0:
        raise Error("mesh_to_volume: empty mesh")
    vs = resolve_voxel_size(V, options.voxel_size)
    lo, dims = band_index_box(V, vs)
    if np.any(dims > (1 << 24)) or float(np.prod(dims.astype(np.float64))) > 2.0**40:
        raise Error(f"Voxel size too small: {vs}")
    own = engine is None
    eng = engine if engine is not None else FastWindingNumber(mesh)
    try:
        origin = (lo.astype(np.float64) * vs).astype(np.float32)
        band = np.float32(EXTERIOR_BANDWIDTH * vs)
        values, active = eng.sdf_grid(origin, (vs, vs, vs), dims, band, signed=options.signing_method == "WindingNumber",
                                      device_output=device_output)
    finally:
        if own:
            eng.close()
    return VolumeGrid(values=values, ijk_min=tuple(int(x) for x in lo), voxel_size=vs, background=float(band), active_voxels=active)
