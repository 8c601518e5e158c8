"""CPU tier: the C-ABI library loads, exports every symbol include/wn_b200.h declares, and refuses to compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "wn_b200.h")).read()
    return sorted(set(re.findall(r"WN_API\s+[\w\s\*]+?\b(wn_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ("wn_create", "wn_create_from_topology", "wn_destroy", "wn_solid_angle", "wn_is_inside", "wn_query_grid", "wn_exact",
                 "wn_tree_pack", "wn_create_from_packed", "wn_last_error"):
        assert must in names
    assert len(names) >= 20


def test_library_exports_every_declared_symbol_and_binding_covers_them():
    from lagrange_b200 import _capi

    L = _capi.lib()
    names = declared_symbols()
    for n in names:
        assert hasattr(L, n), f"{n} declared in wn_b200.h but not exported by libwn_b200.so"
    assert sorted(_capi.SIGNATURES) == names
    assert b"sm_100a" in L.wn_version()


def test_struct_layouts_match_the_header():
    from lagrange_b200 import _capi

    opt = _capi.wn_options()
    assert _capi.lib().wn_options_init(ctypes.byref(opt)) == 0
    assert opt.struct_size == ctypes.sizeof(_capi.wn_options) and opt.accuracy_scale == 2.0 and opt.order == 2
    assert opt.leaf_size == 1 and opt.morton_bits == 63 and opt.device == -1


def test_invalid_arguments_are_reported_not_crashed():
    from lagrange_b200 import _capi

    L = _capi.lib()
    h = ctypes.c_void_p()
    assert L.wn_create(None, 3, None, 1, None, ctypes.byref(h)) == 1
    assert b"null" in L.wn_last_error()
    assert L.wn_solid_angle(None, None, 0, 2.0, 0, None, None) != 0
    assert L.wn_destroy(None) == 0


def test_no_cpu_fallback():
    """Without a CUDA device the engine must fail loudly (status WN_ERR_CUDA), never compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is exercised on CPU-only machines")
    import lagrange_b200 as lb

    V, F = lb.primitive.generate_icosahedron()
    with pytest.raises(lb.Error, match="no CUDA device"):
        lb.FastWindingNumber(V, F)
    with pytest.raises(lb.Error, match="empty"):
        lb.FastWindingNumber().is_inside([0, 0, 0])


def test_surface_mesh_shim_and_constructor_errors():
    import lagrange_b200 as lb

    m = lb.SurfaceMesh(2)
    m.add_vertices(np.zeros((3, 2)))
    with pytest.raises(lb.Error, match="only supports 3D meshes"):
        lb.FastWindingNumber(m)
    m = lb.SurfaceMesh(3)
    m.add_vertices(np.zeros((4, 3)))
    m.add_polygon([0, 1, 2, 3])
    assert not m.is_triangle_mesh()
    with pytest.raises(lb.Error, match="only supports triangle meshes"):
        lb.FastWindingNumber(m)
    m = lb.SurfaceMesh.from_arrays(np.zeros((3, 3), np.float64), np.array([[0, 1, 2]], np.uint64))
    assert m.is_triangle_mesh() and m.get_num_facets() == 1 and m.get_dimension() == 3
