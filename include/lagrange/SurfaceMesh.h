// Minimal stand-in for lagrange::SurfaceMesh<Scalar, Index>: just the members FastWindingNumber's constructor touches
// (adobe/lagrange modules/core/include/lagrange/SurfaceMesh.h: add_vertex :238, add_vertices :259, add_triangle :279,
// add_triangles :300, add_polygon :370, get_dimension :1918, is_triangle_mesh :632, get_num_vertices/facets) plus the
// contiguous row-major coordinate / index buffers that vertex_view / facet_view expose
// (modules/core/src/views.cpp:156-175). When the engine is dropped into a real lagrange checkout this header is not
// used: the real SurfaceMesh and views are.
#pragma once

#include <array>
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <vector>

namespace lagrange {

/// Counterpart of lagrange::Error (modules/core/include/lagrange/utils/Error.h): what la_runtime_assert throws.
struct Error : public std::runtime_error
{
    using std::runtime_error::runtime_error;
};

template <typename Scalar_, typename Index_>
class SurfaceMesh
{
public:
    using Scalar = Scalar_;
    using Index = Index_;

    explicit SurfaceMesh(Index dimension = 3)
        : m_dim(dimension)
    {}

    void add_vertex(std::initializer_list<Scalar> p)
    {
        if (p.size() != static_cast<size_t>(m_dim)) throw Error("add_vertex: wrong dimension");
        m_vertices.insert(m_vertices.end(), p.begin(), p.end());
    }
    /// Appends `num_vertices` points read from `coordinates` (row-major, get_dimension() scalars each).
    void add_vertices(Index num_vertices, const Scalar* coordinates)
    {
        m_vertices.insert(m_vertices.end(), coordinates, coordinates + static_cast<size_t>(num_vertices) * m_dim);
    }
    void add_triangle(Index v0, Index v1, Index v2) { add_polygon({v0, v1, v2}); }
    void add_triangles(Index num_facets, const Index* indices)
    {
        for (Index f = 0; f < num_facets; ++f) add_triangle(indices[3 * f], indices[3 * f + 1], indices[3 * f + 2]);
    }
    void add_polygon(std::initializer_list<Index> corners)
    {
        m_corners.insert(m_corners.end(), corners.begin(), corners.end());
        m_facet_end.push_back(static_cast<Index>(m_corners.size()));
    }

    Index get_dimension() const { return m_dim; }
    Index get_num_vertices() const { return static_cast<Index>(m_vertices.size() / m_dim); }
    Index get_num_facets() const { return static_cast<Index>(m_facet_end.size()); }
    Index get_facet_size(Index f) const { return m_facet_end[f] - (f == 0 ? Index(0) : m_facet_end[f - 1]); }

    /// True when every facet has 3 corners. An empty mesh qualifies, as in core/src/SurfaceMesh.cpp:2320-2323.
    bool is_triangle_mesh() const
    {
        for (Index f = 0; f < get_num_facets(); ++f)
            if (get_facet_size(f) != 3) return false;
        return true;
    }

    /// Row-major V (num_vertices x dimension) and corner list; for a triangle mesh the latter is F (num_facets x 3).
    const Scalar* vertex_data() const { return m_vertices.data(); }
    const Index* corner_data() const { return m_corners.data(); }

private:
    Index m_dim;
    std::vector<Scalar> m_vertices;
    std::vector<Index> m_corners;
    std::vector<Index> m_facet_end;
};

} // namespace lagrange
