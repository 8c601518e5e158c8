// lagrange::winding::FastWindingNumber — B200 engine behind the reference's public surface.
//
// Drop-in for adobe/lagrange modules/winding/include/lagrange/winding/FastWindingNumber.h:24-108. The declarations a
// caller of the reference uses are unchanged:
//     template <Scalar, Index> FastWindingNumber(const SurfaceMesh<Scalar, Index>&)   (:40)
//     FastWindingNumber(); ~FastWindingNumber(); move ctor / move assign noexcept; copies deleted   (:45-82)
//     bool  is_inside(const std::array<float, 3>&) const                               (:91)
//     float solid_angle(const std::array<float, 3>&) const                             (:100)
// What is behind them is different: the PIMPL owns an opaque handle of the CUDA engine (include/wn_b200.h) instead of
// an HDK_Sample::UT_SolidAngle. Added on top (BASELINE.json north_star): batched overloads on spans of
// points, implicit cell-centred lattices, an accuracy parameter, an exact brute-force mode and build options.
//
// Error behaviour follows the reference: non-3D or non-triangle input throws lagrange::Error with the reference's two
// messages (modules/winding/src/FastWindingNumber.cpp:91-96). Querying a default-constructed engine throws instead
// of dereferencing a null PIMPL.
#pragma once

#include <lagrange/SurfaceMesh.h>

#include <array>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

struct wn_engine;

namespace lagrange {
namespace winding {

/// Hierarchy built on the GPU (wn_hierarchy of the C-ABI, include/wn_b200.h).
enum class Hierarchy : int {
    Reference = 3, ///< the reference builder's own 4-ary SAH tree (UT_BVH<4>), reproduced on the GPU: results match the
                   ///< reference algorithm to float rounding. Default: this class is a drop-in.
    Lbvh = 0, ///< Morton codes + Karras: fastest build, ~20 % slower queries, results ~1e-3 from the reference's
    Kd = 1, ///< balanced k-d (object-median splits)
    KdSah = 2, ///< k-d with binned-SAH cuts and multi-triangle leaves: slightly faster queries than Reference
};

/// Build / query options that the reference hard-codes (UT_SolidAngle defaults: order 2, accuracy scale 2).
struct FastWindingNumberOptions
{
    float accuracy_scale = 2.f; ///< beta: a cluster is expanded when |q - P| > beta * R
    int order = 2; ///< Taylor order of the far field (0, 1, 2)
    Hierarchy hierarchy = Hierarchy::Reference;
    int leaf_size = 1; ///< max triangles per leaf (1..16); ignored by Hierarchy::Reference (one triangle per leaf slot)
    int morton_bits = 63; ///< Hierarchy::Lbvh: 30 or 63
    bool vertex_radius = false; ///< exact cluster radius instead of the reference's box-corner bound (changes results)
    bool balanced_hierarchy = false; ///< deprecated spelling of hierarchy = Hierarchy::Kd
    int device = -1; ///< CUDA device, -1 = current
    /// More than one entry: the tree is built on devices[0] and replicated to the others (wn_replicate), and the whole-lattice
    /// overloads shard their tile layers over all of them from one host thread per GPU (wn_query_grid_multi). Overrides `device`.
    std::vector<int> devices;
    /// The reference-surface single-point overloads is_inside(pos) / solid_angle(pos) are called once per voxel from many
    /// host threads (modules/volume/src/mesh_to_volume.cpp:175-183). true (default): they walk a HOST copy of the packed
    /// tree (made on first use, lock-free afterwards, ~1-3 us per call and thread). false: every call is a kernel launch
    /// (~85 us, serialised per engine). Batched overloads always run on the GPU.
    bool host_single_point = true;
};

/// Lattice p(i,j,k) = origin + spacing * ((i,j,k) + 1/2), x fastest: the samples mesh_to_volume evaluates
/// (modules/volume/src/mesh_to_volume.cpp:147-149,175-182).
struct Lattice
{
    std::array<float, 3> origin{};
    std::array<float, 3> spacing{};
    std::array<int64_t, 3> dims{};
};

class FastWindingNumber
{
public:
    template <typename Scalar, typename Index>
    FastWindingNumber(const SurfaceMesh<Scalar, Index>& mesh);

    template <typename Scalar, typename Index>
    FastWindingNumber(const SurfaceMesh<Scalar, Index>& mesh, const FastWindingNumberOptions& options);

    FastWindingNumber();
    ~FastWindingNumber();
    FastWindingNumber(FastWindingNumber&& other) noexcept;
    FastWindingNumber& operator=(FastWindingNumber&& other) noexcept;
    FastWindingNumber(const FastWindingNumber& other) = delete;
    FastWindingNumber& operator=(const FastWindingNumber& other) = delete;

    /// (double)solid_angle(pos) / (4 pi) > 0.5, the reference's predicate.
    bool is_inside(const std::array<float, 3>& pos) const;
    /// Solid angle subtended by the mesh at pos.
    float solid_angle(const std::array<float, 3>& pos) const;

    // ---- batched overloads (host or device pointers; n points as packed xyz triples) --------------------------------
    void is_inside(const float* xyz, size_t n, uint8_t* out) const;
    void solid_angle(const float* xyz, size_t n, float* out) const;
    /// Whole lattice or the z-slab [z_begin, z_end); outputs hold dims[0]*dims[1]*(z_end - z_begin) values.
    void is_inside(const Lattice& lattice, uint8_t* out, int64_t z_begin = 0, int64_t z_end = -1) const;
    void solid_angle(const Lattice& lattice, float* out, int64_t z_begin = 0, int64_t z_end = -1) const;
    /// Same classification, one BIT per lattice point: point i -> bit (i & 7) of out[i >> 3]; out holds
    /// (dims[0]*dims[1]*(z_end - z_begin) + 7) / 8 bytes. An eighth of the device-to-host traffic of the byte overload.
    void is_inside_bits(const Lattice& lattice, uint8_t* out, int64_t z_begin = 0, int64_t z_end = -1) const;
    /// Narrow-band signed distance at the lattice's cell centres: min(distance to the mesh, band), negative where is_inside
    /// holds; what volume::mesh_to_volume asks of OpenVDB with Sign::WindingNumber (mesh_to_volume.cpp:160-183, band = 3 voxels).
    /// Returns the number of cells with distance < band. signed_distance = false skips the predicate.
    int64_t signed_distance(const Lattice& lattice, float band, float* out, bool signed_distance = true) const;
    /// Sparse variant: only the cells of the band (|d| < band), ordered by linear index (z*ny + y)*nx + x -- what an OpenVDB grid keeps
    /// active. Returns the size of the band; fills at most `capacity` cells (call with capacity 0 to size the buffers).
    int64_t signed_distance_sparse(const Lattice& lattice, float band, int64_t capacity, int64_t* index, float* value,
                                   bool signed_distance = true) const;
    /// Closest point of the mesh for n query points (bvh::TriangleAABBTree::get_closest_point, modules/bvh/include/lagrange/bvh/
    /// TriangleAABBTree.h:84-88): squared distance, triangle id, closest point; any output may be null (not all).
    void closest_point(const float* xyz, size_t n, float* sq_dist, int32_t* triangle, float* point, float max_distance = 0.f) const;
    /// Exact mode: brute-force sum over all triangles (no hierarchy, no approximation).
    void exact_solid_angle(const float* xyz, size_t n, float* out) const;

    /// beta used by the queries above (initially options.accuracy_scale).
    void set_accuracy_scale(float beta);
    float accuracy_scale() const;
    /// Device time of the hierarchy build in milliseconds, number of triangles, bytes of the packed tree.
    float build_milliseconds() const;
    int64_t num_triangles() const;
    int64_t tree_bytes() const;

protected:
    struct Impl;
    std::unique_ptr<Impl> m_impl; // move-only like the reference's value_ptr<Impl>; moved-from engines are empty

private:
    void initialize(const float* vertices, int64_t num_vertices, const int32_t* triangles, int64_t num_triangles,
                    const FastWindingNumberOptions& options);
    const wn_engine* engine() const;
};

} // namespace winding
} // namespace lagrange
