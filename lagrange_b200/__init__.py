"""lagrange_b200 — B200-native (sm_100a) fast generalized winding numbers behind lagrange's FastWindingNumber surface.

    from lagrange_b200 import FastWindingNumber, SurfaceMesh, primitive

The compute path is libwn_b200.so (hand-written CUDA, C-ABI in include/wn_b200.h). Importing this package does not
need a GPU; constructing an engine does, and fails loudly without one (no CPU fallback).
"""
from . import callers, io, primitive, volume
from .mesh import SurfaceMesh
from .winding import Error, FastWindingNumber

__all__ = ["FastWindingNumber", "SurfaceMesh", "Error", "primitive", "callers", "volume", "io"]
__version__ = "0.1.0"
