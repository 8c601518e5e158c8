// Host layer of lagrange::winding::FastWindingNumber over the C-ABI of the CUDA engine (include/wn_b200.h).
// Mirrors the structure of adobe/lagrange modules/winding/src/FastWindingNumber.cpp: an Impl that converts the mesh to
// float / int (there :40-52), hands it to the engine (there m_engine.init, :57; here wn_create) and forwards queries
// (there computeSolidAngle, :66,75; here wn_is_inside / wn_solid_angle), with the same explicit instantiations over
// (float|double) x (uint32|uint64) (there :116-118 via LA_SURFACE_MESH_X).
#include <lagrange/winding/FastWindingNumber.h>

#include <wn_b200.h>

// The single-point overloads of the reference surface walk a host copy of the packed tree with the engine's own per-point
// traversal (wn_traverse_point: the definition the warp kernels mirror lane-wise), compiled here for the host. This is the
// library's code, not the test oracle; everything batched runs on the GPU and fails loudly without one.
#include "../lagrange_b200/csrc/wn_device.cuh"
#include "../lagrange_b200/csrc/wn_packed.h"

#include <limits>
#include <mutex>
#include <string>
#include <vector>

namespace lagrange {
namespace winding {

namespace {
void check(wn_status s)
{
    if (s != WN_OK) throw Error(std::string("FastWindingNumber: ") + wn_last_error());
}
} // namespace

struct FastWindingNumber::Impl
{
    wn_engine* engine = nullptr;
    float beta = 2.f;
    bool host_single_point = true;
    // host copy of the packed tree for the single-point overloads: made once, on first use, read-only afterwards
    mutable std::once_flag host_once;
    mutable std::vector<char> host_blob;
    mutable WnTreeView host_view{};
    mutable std::string host_error;
    std::vector<wn_engine*> all; // engine + its replicas on other GPUs (options.devices); empty when single-GPU
    ~Impl()
    {
        for (size_t i = 1; i < all.size(); ++i) wn_destroy(all[i]);
        wn_destroy(engine);
    }

    const WnTreeView* host_tree() const
    {
        std::call_once(host_once, [this]() {
            int64_t nbytes = 0;
            if (wn_tree_packed_size(engine, &nbytes) != WN_OK || nbytes < (int64_t)sizeof(WnPackedHeader)) {
                host_error = wn_last_error();
                return;
            }
            host_blob.resize(static_cast<size_t>(nbytes) + 64);
            char* base = host_blob.data() + (64 - (reinterpret_cast<uintptr_t>(host_blob.data()) & 63)) % 64; // float4 loads want 16-byte alignment
            if (wn_tree_pack(engine, base, nbytes, nullptr) != WN_OK) {
                host_error = wn_last_error();
                host_blob.clear();
                return;
            }
            WnPackedHeader h;
            memcpy(&h, base, sizeof(h));
            if (h.magic != WN_PACKED_MAGIC || h.total_bytes > nbytes) {
                host_error = "packed tree header mismatch";
                host_blob.clear();
                return;
            }
            host_view.hot = reinterpret_cast<const float4*>(base + h.off_hot);
            host_view.cold = reinterpret_cast<const float4*>(base + h.off_cold);
            host_view.kids = reinterpret_cast<const int4*>(base + h.off_kids);
            host_view.tri = reinterpret_cast<const float4*>(base + h.off_tris);
            host_view.n_entries = static_cast<int>(h.n_entries);
            host_view.n_tris = static_cast<int>(h.n_tris);
        });
        if (host_blob.empty()) throw Error(std::string("FastWindingNumber: cannot copy the tree to the host: ") + host_error);
        return &host_view;
    }
    float host_solid_angle(const std::array<float, 3>& pos) const
    {
        const WnTreeView* t = host_tree();
        return wn_traverse_point(*t, pos[0], pos[1], pos[2], beta * beta, nullptr);
    }
};

void FastWindingNumber::initialize(const float* vertices, int64_t num_vertices, const int32_t* triangles, int64_t num_triangles,
                                   const FastWindingNumberOptions& options)
{
    wn_options opt;
    check(wn_options_init(&opt));
    opt.accuracy_scale = options.accuracy_scale;
    opt.order = options.order;
    opt.leaf_size = options.leaf_size;
    opt.morton_bits = options.morton_bits;
    opt.hierarchy = options.balanced_hierarchy ? WN_HIERARCHY_KD : static_cast<int>(options.hierarchy);
    opt.radius_mode = options.vertex_radius ? WN_RADIUS_VERTEX : WN_RADIUS_BOX_CORNER;
    opt.device = options.devices.empty() ? options.device : options.devices[0];
    m_impl = std::make_unique<Impl>();
    m_impl->beta = options.accuracy_scale;
    m_impl->host_single_point = options.host_single_point;
    check(wn_create(vertices, num_vertices, triangles, num_triangles, &opt, &m_impl->engine));
    if (options.devices.size() > 1) {
        std::vector<int32_t> others(options.devices.begin() + 1, options.devices.end());
        std::vector<wn_engine*> reps(others.size(), nullptr);
        check(wn_replicate(m_impl->engine, others.data(), static_cast<int32_t>(others.size()), reps.data()));
        m_impl->all.push_back(m_impl->engine);
        m_impl->all.insert(m_impl->all.end(), reps.begin(), reps.end());
    }
}

template <typename Scalar, typename Index>
FastWindingNumber::FastWindingNumber(const SurfaceMesh<Scalar, Index>& mesh)
    : FastWindingNumber(mesh, FastWindingNumberOptions{})
{}

template <typename Scalar, typename Index>
FastWindingNumber::FastWindingNumber(const SurfaceMesh<Scalar, Index>& mesh, const FastWindingNumberOptions& options)
{
    if (mesh.get_dimension() != 3) throw Error("Fast winding number engine only supports 3D meshes");
    if (!mesh.is_triangle_mesh()) throw Error("Fast winding number engine only supports triangle meshes");
    const size_t nv = static_cast<size_t>(mesh.get_num_vertices());
    const size_t nf = static_cast<size_t>(mesh.get_num_facets());
    if (nf > static_cast<size_t>(std::numeric_limits<int32_t>::max() / 3) || nv > static_cast<size_t>(std::numeric_limits<int32_t>::max()))
        throw Error("Fast winding number engine: mesh too large for 32-bit indices");
    // the engine works on float coordinates and int indices, whatever the mesh stores
    std::vector<float> vertices(nv * 3);
    const Scalar* vsrc = mesh.vertex_data();
    for (size_t i = 0; i < nv * 3; ++i) vertices[i] = static_cast<float>(vsrc[i]);
    std::vector<int32_t> triangles(nf * 3);
    const Index* fsrc = mesh.corner_data();
    for (size_t i = 0; i < nf * 3; ++i) triangles[i] = static_cast<int32_t>(fsrc[i]);
    initialize(vertices.data(), static_cast<int64_t>(nv), triangles.data(), static_cast<int64_t>(nf), options);
}

FastWindingNumber::FastWindingNumber() = default;
FastWindingNumber::~FastWindingNumber() = default;
FastWindingNumber::FastWindingNumber(FastWindingNumber&& other) noexcept = default;
FastWindingNumber& FastWindingNumber::operator=(FastWindingNumber&& other) noexcept = default;

const wn_engine* FastWindingNumber::engine() const
{
    if (!m_impl || !m_impl->engine) throw Error("FastWindingNumber: engine is empty (default constructed or moved from)");
    return m_impl->engine;
}

bool FastWindingNumber::is_inside(const std::array<float, 3>& pos) const
{
    const wn_engine* e = engine(); // throws on an empty engine before m_impl is touched
    if (m_impl->host_single_point) return wn_inside_from_omega(m_impl->host_solid_angle(pos));
    uint8_t r = 0;
    check(wn_is_inside(e, pos.data(), 1, m_impl->beta, WN_QUERY_PRESORTED, &r, nullptr));
    return r != 0;
}

float FastWindingNumber::solid_angle(const std::array<float, 3>& pos) const
{
    const wn_engine* e = engine();
    if (m_impl->host_single_point) return m_impl->host_solid_angle(pos);
    float r = 0;
    check(wn_solid_angle(e, pos.data(), 1, m_impl->beta, WN_QUERY_PRESORTED, &r, nullptr));
    return r;
}

void FastWindingNumber::is_inside(const float* xyz, size_t n, uint8_t* out) const
{
    const wn_engine* e = engine();
    check(wn_is_inside(e, xyz, static_cast<int64_t>(n), m_impl->beta, WN_QUERY_DEFAULT, out, nullptr));
}

void FastWindingNumber::solid_angle(const float* xyz, size_t n, float* out) const
{
    const wn_engine* e = engine();
    check(wn_solid_angle(e, xyz, static_cast<int64_t>(n), m_impl->beta, WN_QUERY_DEFAULT, out, nullptr));
}

void FastWindingNumber::is_inside(const Lattice& l, uint8_t* out, int64_t z_begin, int64_t z_end) const
{
    const wn_engine* e = engine();
    if (m_impl->all.size() > 1 && z_begin == 0 && (z_end < 0 || z_end == l.dims[2])) {
        check(wn_query_grid_multi(m_impl->all.data(), static_cast<int32_t>(m_impl->all.size()), l.origin.data(), l.spacing.data(), l.dims.data(),
                                  m_impl->beta, WN_QUERY_DEFAULT, nullptr, out));
        return;
    }
    check(wn_query_grid(e, l.origin.data(), l.spacing.data(), l.dims.data(), z_begin, z_end < 0 ? l.dims[2] : z_end, m_impl->beta, WN_QUERY_DEFAULT, nullptr,
                        out, nullptr));
}

void FastWindingNumber::is_inside_bits(const Lattice& l, uint8_t* out, int64_t z_begin, int64_t z_end) const
{
    const wn_engine* e = engine();
    if (m_impl->all.size() > 1 && z_begin == 0 && (z_end < 0 || z_end == l.dims[2])) {
        check(wn_query_grid_multi(m_impl->all.data(), static_cast<int32_t>(m_impl->all.size()), l.origin.data(), l.spacing.data(), l.dims.data(),
                                  m_impl->beta, WN_QUERY_OUT_BITS, nullptr, out));
        return;
    }
    check(wn_query_grid(e, l.origin.data(), l.spacing.data(), l.dims.data(), z_begin, z_end < 0 ? l.dims[2] : z_end, m_impl->beta, WN_QUERY_OUT_BITS,
                        nullptr, out, nullptr));
}

void FastWindingNumber::solid_angle(const Lattice& l, float* out, int64_t z_begin, int64_t z_end) const
{
    const wn_engine* e = engine();
    if (m_impl->all.size() > 1 && z_begin == 0 && (z_end < 0 || z_end == l.dims[2])) {
        check(wn_query_grid_multi(m_impl->all.data(), static_cast<int32_t>(m_impl->all.size()), l.origin.data(), l.spacing.data(), l.dims.data(),
                                  m_impl->beta, WN_QUERY_DEFAULT, out, nullptr));
        return;
    }
    check(wn_query_grid(e, l.origin.data(), l.spacing.data(), l.dims.data(), z_begin, z_end < 0 ? l.dims[2] : z_end, m_impl->beta, WN_QUERY_DEFAULT, out,
                        nullptr, nullptr));
}

int64_t FastWindingNumber::signed_distance(const Lattice& l, float band, float* out, bool signed_distance) const
{
    const wn_engine* e = engine();
    int64_t active = 0;
    check(wn_sdf_grid(e, l.origin.data(), l.spacing.data(), l.dims.data(), band, m_impl->beta, signed_distance ? 0u : WN_SDF_UNSIGNED, out, &active,
                      nullptr));
    return active;
}

int64_t FastWindingNumber::signed_distance_sparse(const Lattice& l, float band, int64_t capacity, int64_t* index, float* value, bool signed_distance) const
{
    const wn_engine* e = engine();
    int64_t active = 0;
    check(wn_sdf_grid_sparse(e, l.origin.data(), l.spacing.data(), l.dims.data(), band, m_impl->beta, signed_distance ? 0u : WN_SDF_UNSIGNED, capacity, index,
                             value, nullptr, &active, nullptr));
    return active;
}

void FastWindingNumber::closest_point(const float* xyz, size_t n, float* sq_dist, int32_t* triangle, float* point, float max_distance) const
{
    check(wn_closest_point(engine(), xyz, static_cast<int64_t>(n), max_distance, WN_QUERY_DEFAULT, sq_dist, triangle, point, nullptr));
}

void FastWindingNumber::exact_solid_angle(const float* xyz, size_t n, float* out) const
{
    check(wn_exact(engine(), xyz, static_cast<int64_t>(n), out, nullptr, nullptr));
}

void FastWindingNumber::set_accuracy_scale(float beta)
{
    if (!(beta > 0.f)) throw Error("FastWindingNumber: accuracy scale must be positive");
    engine();
    m_impl->beta = beta;
}

float FastWindingNumber::accuracy_scale() const
{
    engine();
    return m_impl->beta;
}

float FastWindingNumber::build_milliseconds() const
{
    wn_info info;
    check(wn_get_info(engine(), &info));
    return info.build_ms;
}

int64_t FastWindingNumber::num_triangles() const
{
    wn_info info;
    check(wn_get_info(engine(), &info));
    return info.num_triangles;
}

int64_t FastWindingNumber::tree_bytes() const
{
    wn_info info;
    check(wn_get_info(engine(), &info));
    return info.tree_bytes;
}

// Same (Scalar, Index) grid as LA_SURFACE_MESH_X (modules/core/include/lagrange/SurfaceMeshTypes.h:49-53).
#define WN_INSTANTIATE(Scalar, Index)                                                        \
    template FastWindingNumber::FastWindingNumber(const SurfaceMesh<Scalar, Index>& mesh); \
    template FastWindingNumber::FastWindingNumber(const SurfaceMesh<Scalar, Index>& mesh, const FastWindingNumberOptions&);
WN_INSTANTIATE(float, uint32_t)
WN_INSTANTIATE(double, uint32_t)
WN_INSTANTIATE(float, uint64_t)
WN_INSTANTIATE(double, uint64_t)
#undef WN_INSTANTIATE

} // namespace winding
} // namespace lagrange
