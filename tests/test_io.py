"""OBJ reader (lagrange_b200/io.py)."""
import numpy as np
import pytest


def test_obj_round_trip_and_corner_syntax(tmp_path, prim):
    from lagrange_b200 import io

    V, F = prim.generate_torus(5, 1, 8, 5)
    path = tmp_path / "torus.obj"
    io.save_obj(path, V, F)
    m = io.load_obj(path)
    assert m.is_triangle_mesh() and m.get_num_vertices() == len(V) and m.get_num_facets() == len(F)
    assert np.array_equal(m.facets.astype(np.int64), F.astype(np.int64))
    assert np.allclose(m.vertices, V, rtol=0, atol=1e-6)
    # v/vt/vn corners, negative indices, comments, a quad
    (tmp_path / "mixed.obj").write_text(
        "# comment\nv 0 0 0\nv 1 0 0\nv 1 1 0\nvt 0 0\nvn 0 0 1\nf 1/1/1 2/1/1 3/1/1\nv 0 1 0\nf -4 -3 -2 -1\nf 1//1 3//1 4//1\n")
    q = io.load_obj(tmp_path / "mixed.obj")
    assert not q.is_triangle_mesh() and q.get_num_facets() == 3
    t = io.load_obj(tmp_path / "mixed.obj", triangulate=True)
    assert t.is_triangle_mesh() and t.get_num_facets() == 4
    assert t.facets.tolist() == [[0, 1, 2], [0, 1, 2], [0, 2, 3], [0, 2, 3]]


def test_obj_errors(tmp_path):
    from lagrange_b200 import io
    from lagrange_b200.winding import Error

    (tmp_path / "bad.obj").write_text("v 0 0 0\nv 1 0 0\nf 1 2 5\n")
    with pytest.raises(Error, match="out of range"):
        io.load_obj(tmp_path / "bad.obj")
    (tmp_path / "short.obj").write_text("v 0 0\n")
    with pytest.raises(Error, match="three coordinates"):
        io.load_obj(tmp_path / "short.obj")
