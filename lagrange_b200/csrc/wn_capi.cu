// wn_capi.cu — the C-ABI of libwn_b200.so (include/wn_b200.h): engine lifetime, build orchestration, query launchers.
// The only translation unit of the library; kernels live in the .cuh files it includes.
//
// Reference call sites replaced (adobe/lagrange, modules/winding/src/FastWindingNumber.cpp):
//   :54-57  m_engine.init(...)                 -> wn_create / wn_create_from_topology
//   :66     computeSolidAngle(q)/(4pi) > 0.5   -> wn_is_inside / wn_query_grid(out_inside)
//   :75     computeSolidAngle(q)               -> wn_solid_angle / wn_query_grid(out_omega)
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/wn_b200.h"
#include "wn_packed.h"
#include "wn_query.cuh"
#include "wn_kd.cuh"
#include "wn_refbuild.cuh"
#include "wn_sdf.cuh"

namespace {

thread_local std::string g_last_error;

wn_status fail(wn_status s, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return s;
}

#define WN_CUDA(call)                                                                                              \
    do {                                                                                                           \
        cudaError_t err__ = (call);                                                                                \
        if (err__ != cudaSuccess) {                                                                                \
            cudaGetLastError();                                                                                    \
            return fail(err__ == cudaErrorMemoryAllocation ? WN_ERR_OUT_OF_MEMORY : WN_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
                        cudaGetErrorString(err__), __FILE__, __LINE__);                                            \
        }                                                                                                          \
    } while (0)

struct DeviceGuard
{
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            cudaGetLastError();
            prev = -1;
            return;
        }
        ok = (dev == prev) || cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct DevBuf
{
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t need)
    {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        const size_t want = need + need / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

struct PinnedBuf
{
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t need)
    {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMallocHost(&p, need);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
    }
};

bool is_device_pointer(const void* p)
{
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

size_t align_up(size_t v, size_t a)
{
    return (v + a - 1) / a * a;
}

using PackedHeader = WnPackedHeader; // wn_packed.h: shared with the C++ host layer (single-point queries on a host copy)
constexpr uint64_t kMagic = WN_PACKED_MAGIC;
PackedHeader make_header(int64_t n_entries, int64_t n_tris)
{
    return wn_packed_make_header(n_entries, n_tris);
}

} // namespace

struct wn_engine
{
    int device = 0;
    wn_options opt;
    wn_info info;
    PackedHeader hdr;
    char* blob = nullptr; // packed tree on the device
    WnTreeView view;
    // kept build data (keep_build_data)
    float4* kept_local = nullptr;
    int* kept_child = nullptr;
    unsigned* kept_prim = nullptr;
    int kept_nI = 0, kept_nL = 0, kept_W = 0;
    std::vector<void*> kept_allocs; // device allocations that back the kept arrays
    // per-engine query scratch, guarded by mu
    mutable std::mutex mu;
    // the per-engine scratch is reused by every call: work submitted on a different stream than the previous call's waits for it
    mutable cudaStream_t last_stream = nullptr;
    mutable cudaEvent_t ev_last = nullptr;
    mutable cudaEvent_t ev_batch = nullptr; // orders a finished batch before its copy on copy_stream
    mutable bool ev_last_valid = false;
    mutable std::mutex sdf_mu; // wn_sdf_grid runs two passes (sign, distance) over shared scratch: one caller at a time
    mutable DevBuf s_in, s_out_f, s_out_b, s_out_bits, s_sort, s_stats, s_partial, s_sdf_inside, s_sdf_dense;
    // plan scratch of the tiled path, one set per lane: consecutive batches of a call alternate between two streams, so that the
    // launch tail of one batch's kernels is filled by the other batch's (see dispatch_query)
    struct PlanScratch
    {
        DevBuf hdr, items, samples, order, lvl;
        void release() { hdr.release(), items.release(), samples.release(), order.release(), lvl.release(); }
    };
    mutable PlanScratch s_plan[2];
    mutable cudaStream_t lane_stream = nullptr; // second lane; the first is the caller's stream
    mutable cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_lane[2] = {nullptr, nullptr};
    mutable PinnedBuf p_small;
    mutable cudaStream_t copy_stream = nullptr; // D2H of finished batches while the next batch computes
    mutable int query_slots = 0;                // CTAs of k_tile_query the device holds at once (SMs x occupancy), set on first use
    mutable int64_t last_plan_tiles = 0;        // tiles of the last k_tile_plan launch (wn_debug_last_plan)
    mutable float last_probe_share = -1.0f;     // far-set share measured by the last tiling probe (diagnostics)
    mutable float probe_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    mutable int probe_dims[3] = {0, 0, 0};
    mutable bool probe_cache_valid = false, probe_cache_tiled = false;
};

namespace {

void set_view(wn_engine* e)
{
    e->view.hot = (const float4*)(e->blob + e->hdr.off_hot);
    e->view.cold = (const float4*)(e->blob + e->hdr.off_cold);
    e->view.kids = (const int4*)(e->blob + e->hdr.off_kids);
    e->view.tri = (const float4*)(e->blob + e->hdr.off_tris);
    e->view.n_entries = (int)e->hdr.n_entries;
    e->view.n_tris = (int)e->hdr.n_tris;
}

wn_status resolve_device(const wn_options* opt, int* dev)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(WN_ERR_CUDA, "no CUDA device available (%s): libwn_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    int d = opt ? opt->device : -1;
    if (d < 0) WN_CUDA(cudaGetDevice(&d));
    if (d >= count) return fail(WN_ERR_INVALID_ARGUMENT, "device %d out of range (%d devices)", d, count);
    *dev = d;
    return WN_OK;
}

struct Timer
{
    cudaEvent_t ev[8];
    int n = 0;
    cudaStream_t s;
    explicit Timer(cudaStream_t st) : s(st)
    {
        for (auto& e : ev) cudaEventCreate(&e);
    }
    ~Timer()
    {
        for (auto& e : ev) cudaEventDestroy(e);
    }
    void mark() { cudaEventRecord(ev[n++], s); }
    float ms(int a, int b)
    {
        float m = 0;
        cudaEventElapsedTime(&m, ev[a], ev[b]);
        return m;
    }
};

// Build scratch comes out of ONE device allocation made before the timed region (cudaMalloc inside it would be what the
// build time measures); anything that does not fit falls back to its own cudaMalloc.
struct BuildArena
{
    char* base = nullptr;
    size_t cap = 0, off = 0, used = 0;
    std::vector<void*>* extra = nullptr;
    cudaError_t take(void** p, size_t bytes)
    {
        const size_t need = align_up(bytes ? bytes : 1, 256);
        used += bytes;
        if (base && off + need <= cap) {
            *p = base + off;
            off += need;
            return cudaSuccess;
        }
        cudaError_t r = cudaMalloc(p, need);
        if (r == cudaSuccess) extra->push_back(*p);
        return r;
    }
};

// Everything after the topology exists: moments, radii, packing. `b` has mesh + topology + options filled in.
wn_status finish_build(wn_engine* e, WnBuild& b, BuildArena& arena, size_t blob_cap, Timer& tm, cudaStream_t st)
{
    const int64_t nN = (int64_t)b.nI + b.nL;
    auto dalloc = [&](void** p, size_t bytes) -> cudaError_t { return arena.take(p, bytes); };
    int* small = nullptr; // err, max_depth
    WN_CUDA(dalloc((void**)&b.local, (size_t)nN * 9 * sizeof(float4)));
    WN_CUDA(dalloc((void**)&b.arrive, (size_t)b.nI * sizeof(int)));
    WN_CUDA(dalloc((void**)&b.ntri, (size_t)nN * sizeof(int)));
    WN_CUDA(dalloc((void**)&b.size, (size_t)nN * sizeof(int)));
    WN_CUDA(dalloc((void**)&b.collapsed, (size_t)b.nI));
    WN_CUDA(dalloc((void**)&b.r2v, (size_t)nN * sizeof(unsigned)));
    WN_CUDA(dalloc((void**)&small, 2 * sizeof(int)));
    b.err = small;
    b.max_depth = small + 1;
    WN_CUDA(cudaMemsetAsync(b.arrive, 0, (size_t)b.nI * sizeof(int), st));
    WN_CUDA(cudaMemsetAsync(b.collapsed, 0, (size_t)b.nI, st));
    WN_CUDA(cudaMemsetAsync(b.r2v, 0, (size_t)nN * sizeof(unsigned), st));
    WN_CUDA(cudaMemsetAsync(small, 0, 2 * sizeof(int), st));
    wn::k_moments_climb<<<wn::grid_for(b.nL), wn::kBuildThreads, 0, st>>>(b);
    if (b.radius_mode == WN_RADIUS_VERTEX) wn::k_vertex_radius<<<wn::grid_for(b.nL), wn::kBuildThreads, 0, st>>>(b);
    WN_CUDA(cudaGetLastError());
    tm.mark(); // moments done
    int h_small[2] = {0, 0}, h_size = 0, h_ntri = 0;
    WN_CUDA(cudaMemcpyAsync(h_small, small, sizeof(h_small), cudaMemcpyDeviceToHost, st));
    WN_CUDA(cudaMemcpyAsync(&h_size, b.size, sizeof(int), cudaMemcpyDeviceToHost, st));
    WN_CUDA(cudaMemcpyAsync(&h_ntri, b.ntri, sizeof(int), cudaMemcpyDeviceToHost, st));
    WN_CUDA(cudaStreamSynchronize(st));
    if (h_small[0] == 3) return fail(WN_ERR_INVALID_ARGUMENT, "triangle references a vertex index outside [0, num_vertices)");
    if (h_small[0] != 0 || h_ntri != b.nL)
        return fail(WN_ERR_INVALID_ARGUMENT, "malformed hierarchy topology (error %d, %d of %d triangles reachable from the root)",
                    h_small[0], h_ntri, b.nL);
    e->hdr = make_header(h_size, b.nL);
    if (!e->blob || blob_cap < (size_t)e->hdr.total_bytes) { // normally pre-allocated for the worst case before the timed region
        if (e->blob) cudaFree(e->blob);
        e->blob = nullptr;
        WN_CUDA(cudaMalloc((void**)&e->blob, (size_t)e->hdr.total_bytes));
    }
    set_view(e);
    {
        // the alignment gaps between the sections travel with the blob (wn_tree_pack, broadcasts): keep them defined
        const int64_t ends[6] = {(int64_t)sizeof(PackedHeader), e->hdr.off_hot + (int64_t)h_size * 32, e->hdr.off_cold + (int64_t)h_size * 64,
                                 e->hdr.off_kids + (int64_t)h_size * 16, e->hdr.off_tris + (int64_t)b.nL * 48, e->hdr.off_tri_order + (int64_t)b.nL * 4};
        const int64_t starts[6] = {e->hdr.off_hot, e->hdr.off_cold, e->hdr.off_kids, e->hdr.off_tris, e->hdr.off_tri_order, e->hdr.total_bytes};
        for (int k = 0; k < 6; ++k)
            if (starts[k] > ends[k]) WN_CUDA(cudaMemsetAsync(e->blob + ends[k], 0, (size_t)(starts[k] - ends[k]), st));
    }
    b.hot = (float4*)(e->blob + e->hdr.off_hot);
    b.cold = (float4*)(e->blob + e->hdr.off_cold);
    b.kids = (int4*)(e->blob + e->hdr.off_kids);
    b.tris = (float4*)(e->blob + e->hdr.off_tris);
    b.tri_order = (unsigned*)(e->blob + e->hdr.off_tri_order);
    wn::k_pack<<<wn::grid_for(nN), wn::kBuildThreads, 0, st>>>(b);
    WN_CUDA(cudaGetLastError());
    tm.mark(); // pack done
    WN_CUDA(cudaMemcpyAsync(h_small, small, sizeof(h_small), cudaMemcpyDeviceToHost, st));
    WN_CUDA(cudaStreamSynchronize(st));
    if (h_small[0] != 0) return fail(WN_ERR_INVALID_ARGUMENT, "malformed hierarchy topology while packing (error %d)", h_small[0]);
    e->hdr.width = b.W;
    e->hdr.order = b.order;
    e->hdr.accuracy_scale = e->opt.accuracy_scale;
    e->hdr.max_depth = h_small[1];
    e->hdr.num_vertices = b.nV;
    e->hdr.num_tree_nodes = nN;
    WN_CUDA(cudaMemcpyAsync(e->blob, &e->hdr, sizeof(PackedHeader), cudaMemcpyHostToDevice, st));
    WN_CUDA(cudaStreamSynchronize(st));
    return WN_OK;
}

void fill_info(wn_engine* e)
{
    wn_info& in = e->info;
    in.struct_size = sizeof(wn_info);
    in.device = e->device;
    in.num_vertices = e->hdr.num_vertices;
    in.num_triangles = e->hdr.n_tris;
    in.num_tree_nodes = e->hdr.num_tree_nodes;
    in.num_entries = e->hdr.n_entries;
    in.tree_bytes = e->hdr.total_bytes;
    in.max_depth = e->hdr.max_depth;
    in.width = e->hdr.width;
    in.accuracy_scale = e->opt.accuracy_scale;
    in.order = e->hdr.order;
}

wn_status validate_options(const wn_options* in, wn_options* out, bool imported)
{
    wn_options_init(out);
    if (imported) out->approximate_single_triangles = 1;
    if (in) {
        if (in->struct_size != sizeof(wn_options)) return fail(WN_ERR_INVALID_ARGUMENT, "wn_options.struct_size mismatch (call wn_options_init)");
        *out = *in;
    }
    if (imported) out->leaf_size = 1;
    if (!(out->accuracy_scale > 0.0f)) return fail(WN_ERR_INVALID_ARGUMENT, "accuracy_scale must be > 0");
    if (out->order < 0 || out->order > 2) return fail(WN_ERR_INVALID_ARGUMENT, "order must be 0, 1 or 2");
    if (out->leaf_size < 1 || out->leaf_size > WN_MAX_LEAF_SIZE) return fail(WN_ERR_INVALID_ARGUMENT, "leaf_size must be in [1, %d]", WN_MAX_LEAF_SIZE);
    if (out->morton_bits != 30 && out->morton_bits != 63) return fail(WN_ERR_INVALID_ARGUMENT, "morton_bits must be 30 or 63");
    if (out->radius_mode != WN_RADIUS_BOX_CORNER && out->radius_mode != WN_RADIUS_VERTEX) return fail(WN_ERR_INVALID_ARGUMENT, "bad radius_mode");
    if (out->hierarchy < WN_HIERARCHY_LBVH || out->hierarchy > WN_HIERARCHY_REFERENCE) return fail(WN_ERR_INVALID_ARGUMENT, "bad hierarchy");
    return WN_OK;
}

int env_int(const char* name, int dflt); // defined with the query dispatch below

wn_status create_impl(const float* v_xyz, int64_t nV, const int32_t* tri, int64_t nT, const int32_t* child_in, int64_t num_nodes,
                      int32_t width, const wn_options* opt_in, wn_engine** out)
{
    if (!out) return fail(WN_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    if (nV < 0 || nT < 0 || (nV > 0 && !v_xyz) || (nT > 0 && !tri)) return fail(WN_ERR_INVALID_ARGUMENT, "null mesh buffers or negative sizes");
    if (nT >= WN_MAX_TRIANGLES || nV > INT_MAX / 4) return fail(WN_ERR_UNSUPPORTED, "at most 2^27-1 triangles are supported");
    if (child_in && (width < 2 || width > WN_MAX_WIDTH || num_nodes < 1 || num_nodes > INT_MAX / 8))
        return fail(WN_ERR_INVALID_ARGUMENT, "topology width must be 2..4 and num_nodes >= 1");
    // WN_HIERARCHY_REFERENCE: the reference builder's 4-ary tree is built on the device (K3R, wn_refbuild.cuh) and then takes
    // the same route as a caller-supplied topology (moments, radii, packing): one triangle per leaf slot, single triangles
    // approximated like the reference does.
    const bool refh = !child_in && nT > 0 && (!opt_in || (opt_in->struct_size == sizeof(wn_options) && opt_in->hierarchy == WN_HIERARCHY_REFERENCE));
    const bool imported = child_in != nullptr || refh;
    if (refh) {
        width = 4;
        num_nodes = std::max<int64_t>(1, nT - 1); // upper bound; the builder reports the actual count
    }
    wn_options opt;
    wn_status s = validate_options(opt_in, &opt, imported);
    if (s != WN_OK) return s;
    if (refh) opt.approximate_single_triangles = 1;
    int dev = 0;
    s = resolve_device(&opt, &dev);
    if (s != WN_OK) return s;
    DeviceGuard guard(dev);
    if (!guard.ok) return fail(WN_ERR_CUDA, "cannot select CUDA device %d", dev);

    wn_engine* e = new (std::nothrow) wn_engine;
    if (!e) return fail(WN_ERR_OUT_OF_MEMORY, "host allocation failed");
    e->device = dev;
    e->opt = opt;
    memset(&e->info, 0, sizeof(e->info));
    memset(&e->hdr, 0, sizeof(e->hdr));
    memset(&e->view, 0, sizeof(e->view));

    std::vector<void*> temps;
    BuildArena arena;
    arena.extra = &temps;
    size_t blob_cap = 0;
    bool keep_arena = false;
    cudaStream_t st = nullptr; // legacy default stream: build is synchronous for the caller
    auto cleanup = [&](wn_status r) {
        for (void* p : temps) cudaFree(p);
        temps.clear();
        if (arena.base && !(keep_arena && r == WN_OK)) cudaFree(arena.base);
        if (r != WN_OK) {
            wn_destroy(e);
            e = nullptr;
        }
        return r;
    };
    auto dalloc = [&](void** p, size_t bytes) -> cudaError_t { return arena.take(p, bytes); };
#define WN_TRY(expr)                          \
    do {                                      \
        wn_status s__ = (expr);               \
        if (s__ != WN_OK) return cleanup(s__); \
    } while (0)
#define WN_CUDA_C(call)                                                                                          \
    do {                                                                                                         \
        cudaError_t err__ = (call);                                                                              \
        if (err__ != cudaSuccess) {                                                                              \
            cudaGetLastError();                                                                                  \
            return cleanup(fail(err__ == cudaErrorMemoryAllocation ? WN_ERR_OUT_OF_MEMORY : WN_ERR_CUDA,        \
                                "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__)); \
        }                                                                                                        \
    } while (0)

    if (nT == 0) {
        // empty engine: Omega = 0 everywhere (documented choice)
        e->hdr = make_header(0, 0);
        e->hdr.width = imported ? width : 2;
        e->hdr.order = opt.order;
        e->hdr.accuracy_scale = opt.accuracy_scale;
        e->hdr.num_vertices = nV;
        WN_CUDA_C(cudaMalloc((void**)&e->blob, (size_t)e->hdr.total_bytes));
        WN_CUDA_C(cudaMemset(e->blob, 0, (size_t)e->hdr.total_bytes));
        WN_CUDA_C(cudaMemcpy(e->blob, &e->hdr, sizeof(PackedHeader), cudaMemcpyHostToDevice));
        set_view(e);
        fill_info(e);
        *out = e;
        return cleanup(WN_OK);
    }

    Timer tm(st);
    WnBuild b;
    memset(&b, 0, sizeof(b));
    b.nV = (int)nV;
    b.nT = (int)nT;
    b.nL = (int)nT;
    b.leaf_size = opt.leaf_size;
    b.order = opt.order;
    b.radius_mode = opt.radius_mode;
    b.approx_single = opt.approximate_single_triangles;

    // One allocation for all build scratch and one (worst-case sized) for the packed tree, both before the first event.
    {
        const int64_t nI_max = imported ? num_nodes : std::max<int64_t>(1, nT - 1);
        const int64_t nN_max = nI_max + nT;
        const int64_t W = imported ? width : 2;
        size_t need = 0;
        auto add = [&](size_t bytes) { need += align_up(bytes ? bytes : 1, 256); };
        add((size_t)nV * 12);
        add((size_t)nT * 12);
        add(64);
        if (!imported) {
            add((size_t)nT * 8);
            add((size_t)nT * 8);
            add((size_t)nT * 4);
            add((size_t)nT * 4);
            add((size_t)wn::sort_scratch_bytes(nT));
            if (opt.hierarchy != WN_HIERARCHY_LBVH) add((size_t)nT * 6 * sizeof(int)), add((size_t)nT);
            if (opt.hierarchy == WN_HIERARCHY_KD_SAH) {
                for (int r = 0; r < 2; ++r) add((size_t)(nT + 1) * 4), add((size_t)nT * 4), add((size_t)nT);
                add((size_t)nT * 4), add((size_t)nT * 4), add((size_t)nT * 4);
                add((size_t)wn::scan_scratch_elems(nT) * 4 + 256);
                add(((size_t)nT / WN_KDX_MIN_SAH + 1) * WN_KDX_ROW * 4);
                add(64);
            }
        } else {
            add((size_t)nI_max * W * 4);
            add((size_t)nN_max * 4);
            add((size_t)nT * 4);
            if (refh) {
                add(wn_ref_layout(nT, (size_t)wn::scan_scratch_elems(nT) * 4 + 256).total);
                add((size_t)wn::sort_scratch_bytes(nT));
                add(((size_t)nT + 2) * 16);
            }
        }
        add((size_t)nI_max * W * 4); // child
        add((size_t)nN_max * 4);     // parent
        add((size_t)nN_max);         // slot
        add((size_t)nN_max * 9 * sizeof(float4));
        add((size_t)nI_max * 4);
        add((size_t)nN_max * 4);
        add((size_t)nN_max * 4);
        add((size_t)nI_max);
        add((size_t)nN_max * 4);
        add(64);
        need += 4096;
        if (cudaMalloc((void**)&arena.base, need) == cudaSuccess)
            arena.cap = need;
        else
            cudaGetLastError(); // fall back to piecemeal allocations
        const PackedHeader worst = make_header(nN_max, nT);
        if (cudaMalloc((void**)&e->blob, (size_t)worst.total_bytes) == cudaSuccess)
            blob_cap = (size_t)worst.total_bytes;
        else
            cudaGetLastError();
    }

    // mesh on the device (copied: the caller may free its buffers when we return)
    float* d_v = nullptr;
    int* d_tri = nullptr;
    WN_CUDA_C(dalloc((void**)&d_v, (size_t)nV * 3 * sizeof(float)));
    WN_CUDA_C(dalloc((void**)&d_tri, (size_t)nT * 3 * sizeof(int)));
    WN_CUDA_C(cudaMemcpyAsync(d_v, v_xyz, (size_t)nV * 3 * sizeof(float), cudaMemcpyDefault, st));
    WN_CUDA_C(cudaMemcpyAsync(d_tri, tri, (size_t)nT * 3 * sizeof(int), cudaMemcpyDefault, st));
    b.v_xyz = d_v;
    b.tri = d_tri;

    int* d_small = nullptr; // bounds[6], err
    WN_CUDA_C(dalloc((void**)&d_small, 8 * sizeof(int)));
    {
        const int init[8] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN, 0, 0};
        WN_CUDA_C(cudaMemcpyAsync(d_small, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    tm.mark(); // 0: start
    unsigned* d_prim = nullptr;
    if (!imported) {
        b.W = 2;
        b.nI = nT >= 2 ? (int)nT - 1 : 1;
        const int64_t nN = (int64_t)b.nI + b.nL;
        uint64_t *k0 = nullptr, *k1 = nullptr;
        unsigned *v0 = nullptr, *v1 = nullptr;
        void* sort_scratch = nullptr;
        WN_CUDA_C(dalloc((void**)&k0, (size_t)nT * sizeof(uint64_t)));
        WN_CUDA_C(dalloc((void**)&k1, (size_t)nT * sizeof(uint64_t)));
        WN_CUDA_C(dalloc((void**)&v0, (size_t)nT * sizeof(unsigned)));
        WN_CUDA_C(dalloc((void**)&v1, (size_t)nT * sizeof(unsigned)));
        WN_CUDA_C(dalloc(&sort_scratch, (size_t)wn::sort_scratch_bytes(nT)));
        const int blocks = std::min(wn::grid_for(nT), 148 * 8);
        wn::k_centroid_bounds<<<blocks, wn::kBuildThreads, 0, st>>>(d_v, d_tri, (int)nT, (int)nV, d_small, d_small + 6);
        WN_CUDA_C(cudaGetLastError());
        {
            // indices are validated by k_centroid_bounds before any kernel dereferences them
            int h_err = 0;
            WN_CUDA_C(cudaMemcpyAsync(&h_err, d_small + 6, sizeof(int), cudaMemcpyDeviceToHost, st));
            WN_CUDA_C(cudaStreamSynchronize(st));
            if (h_err) return cleanup(fail(WN_ERR_INVALID_ARGUMENT, "triangle references a vertex index outside [0, num_vertices)"));
        }
        const bool kd = opt.hierarchy == WN_HIERARCHY_KD && nT >= 2;
        const bool kdx = opt.hierarchy == WN_HIERARCHY_KD_SAH && nT >= 2;
        const uint64_t* keys = nullptr;
        WN_CUDA_C(dalloc((void**)&b.child, (size_t)b.nI * 2 * sizeof(int)));
        WN_CUDA_C(dalloc((void**)&b.parent, (size_t)nN * sizeof(int)));
        WN_CUDA_C(dalloc((void**)&b.slot, (size_t)nN));
        WN_CUDA_C(cudaMemsetAsync(b.parent, 0xff, (size_t)nN * sizeof(int), st));
        WN_CUDA_C(cudaMemsetAsync(b.slot, 0, (size_t)nN, st));
        if (kdx) {
            // K3'': level-synchronous k-d build with SAH-guided split positions (wn_kd.cuh)
            const int N = (int)nT, leaf = std::max(1, opt.leaf_size);
            int* d_bounds = nullptr;
            unsigned *node_of = nullptr, *lstart[2] = {nullptr, nullptr};
            int *lpid[2] = {nullptr, nullptr}, *d_nl = nullptr, *d_segbox = nullptr, *d_res = nullptr;
            unsigned char *lmeta[2] = {nullptr, nullptr}, *d_skip = nullptr;
            uint32_t *d_cnt = nullptr, *d_scan = nullptr;
            WN_CUDA_C(dalloc((void**)&d_bounds, (size_t)nT * 6 * sizeof(int)));
            WN_CUDA_C(dalloc((void**)&d_skip, (size_t)nT));
            for (int r = 0; r < 2; ++r) {
                WN_CUDA_C(dalloc((void**)&lstart[r], (size_t)(nT + 1) * 4));
                WN_CUDA_C(dalloc((void**)&lpid[r], (size_t)nT * 4));
                WN_CUDA_C(dalloc((void**)&lmeta[r], (size_t)nT));
            }
            WN_CUDA_C(dalloc((void**)&node_of, (size_t)nT * 4));
            WN_CUDA_C(dalloc((void**)&d_nl, (size_t)nT * 4));
            WN_CUDA_C(dalloc((void**)&d_cnt, (size_t)nT * 4));
            WN_CUDA_C(dalloc((void**)&d_scan, (size_t)wn::scan_scratch_elems(nT) * 4 + 256));
            const int64_t seg_rows = nT / WN_KDX_MIN_SAH + 1;
            WN_CUDA_C(dalloc((void**)&d_segbox, (size_t)seg_rows * WN_KDX_ROW * 4));
            WN_CUDA_C(dalloc((void**)&d_res, 64));
            wn::k_iota<<<wn::grid_for(nT), wn::kBuildThreads, 0, st>>>(v0, N);
            WN_CUDA_C(cudaMemsetAsync(node_of, 0, (size_t)nT * 4, st));
            WN_CUDA_C(cudaMemsetAsync(d_skip, 0, (size_t)nT, st));
            {
                const unsigned h_start[2] = {0u, (unsigned)N};
                const int h_pid = -1;
                WN_CUDA_C(cudaMemcpyAsync(lstart[0], h_start, sizeof(h_start), cudaMemcpyHostToDevice, st));
                WN_CUDA_C(cudaMemcpyAsync(lpid[0], &h_pid, sizeof(h_pid), cudaMemcpyHostToDevice, st));
                WN_CUDA_C(cudaMemsetAsync(lmeta[0], 0, 1, st));
                WN_CUDA_C(cudaMemsetAsync(d_res, 0, 64, st));
            }
            tm.mark(); // 1
            unsigned *cur = v0, *other = v1;
            int count = 1, max_range = N;
            for (int level = 0;; ++level) {
                if (level > 160) return cleanup(fail(WN_ERR_CUDA, "k-d build did not terminate"));
                const int r = level & 1;
                wn::KdxLevel L{lstart[r], lpid[r], lmeta[r], count};
                wn::k_kd_init_bounds<<<wn::grid_for((int64_t)count * 6), wn::kBuildThreads, 0, st>>>(d_bounds, count);
                wn::k_kdx_bounds<<<wn::grid_for(nT), wn::kBuildThreads, 0, st>>>(d_v, d_tri, cur, node_of, N, lstart[r], leaf, d_bounds);
                wn::k_kdx_keys<<<wn::grid_for(nT), wn::kBuildThreads, 0, st>>>(d_v, d_tri, cur, node_of, N, lstart[r], leaf, d_bounds, k0);
                int nbits = 0;
                while (((int64_t)1 << nbits) < count) ++nbits;
                const int which = wn::radix_sort_pairs<uint64_t>(k0, cur, k1, other, nT, 0, 16 + nbits, sort_scratch, st);
                if (which) std::swap(cur, other);
                if (max_range >= WN_KDX_MIN_SAH && max_range > leaf) { // otherwise every split of this level is a median split
                    wn::k_kdx_init_segbox_nodes<<<wn::grid_for((int64_t)count * WN_KDX_ROW), wn::kBuildThreads, 0, st>>>(lstart[r], count, leaf, d_segbox);
                    wn::k_kdx_segboxes<<<wn::grid_for(nT), wn::kBuildThreads, 0, st>>>(d_v, d_tri, cur, node_of, N, lstart[r], leaf, d_bounds, d_segbox);
                }
                WN_CUDA_C(cudaMemsetAsync(d_res, 0, 2 * sizeof(int), st));
                WN_CUDA_C(cudaMemsetAsync(d_res + 3, 0, sizeof(int), st));
                wn::k_kdx_split<<<wn::grid_for(count), wn::kBuildThreads, 0, st>>>(L, N, leaf, level, d_segbox, d_nl, d_cnt, d_res);
                wn::exclusive_scan_u32(d_cnt, count, d_scan, st);
                wn::k_kdx_scatter<<<wn::grid_for(count), wn::kBuildThreads, 0, st>>>(L, N, level, d_nl, d_cnt, d_res, lstart[r ^ 1], lpid[r ^ 1],
                                                                                       lmeta[r ^ 1], b.child, b.parent, b.slot, d_skip, d_res + 1);
                wn::k_kdx_assign<<<wn::grid_for(nT), wn::kBuildThreads, 0, st>>>(node_of, N, lstart[r], d_nl, d_cnt);
                int h_res[4] = {0, 0, 0, 0};
                WN_CUDA_C(cudaMemcpyAsync(h_res, d_res, sizeof(h_res), cudaMemcpyDeviceToHost, st));
                WN_CUDA_C(cudaStreamSynchronize(st));
                WN_CUDA_C(cudaGetLastError());
                if (h_res[0] == 0) break; // nothing was split: every range is a finished leaf range, all linked by this pass
                count = h_res[1];
                max_range = h_res[3];
            }
            d_prim = cur;
            b.skip = env_int("WN_KD_WIDE", 1) ? d_skip : nullptr;
            tm.mark(); // 2: order + hierarchy
        } else if (!kd) {
            const int bpa = opt.morton_bits == 63 ? 21 : 10;
            wn::k_morton<uint64_t><<<wn::grid_for(nT), wn::kBuildThreads, 0, st>>>(d_v, d_tri, (int)nT, d_small, bpa, k0, v0);
            WN_CUDA_C(cudaGetLastError());
            tm.mark(); // 1: morton
            const int which = wn::radix_sort_pairs<uint64_t>(k0, v0, k1, v1, nT, 0, opt.morton_bits, sort_scratch, st);
            WN_CUDA_C(cudaGetLastError());
            keys = which ? k1 : k0;
            d_prim = which ? v1 : v0;
            tm.mark(); // 2: sort
        } else {
            // K3': balanced k-d order (wn_kd.cuh): per level, node centroid bounds -> keys -> one stable sort
            const int levels = wn_kd_levels((int)nT, opt.leaf_size);
            int* d_bounds = nullptr;
            WN_CUDA_C(dalloc((void**)&d_bounds, (size_t)nT * 6 * sizeof(int)));
            wn::k_iota<<<wn::grid_for(nT), wn::kBuildThreads, 0, st>>>(v0, (int)nT);
            tm.mark(); // 1
            unsigned *cur = v0, *other = v1;
            for (int l = 0; l < levels; ++l) {
                const int nodes = 1 << l; // < nT
                wn::k_kd_init_bounds<<<wn::grid_for((int64_t)nodes * 6), wn::kBuildThreads, 0, st>>>(d_bounds, nodes);
                wn::k_kd_bounds<<<wn::grid_for(nT), wn::kBuildThreads, 0, st>>>(d_v, d_tri, cur, (int)nT, l, d_bounds);
                wn::k_kd_keys<<<wn::grid_for(nT), wn::kBuildThreads, 0, st>>>(d_v, d_tri, cur, (int)nT, l, d_bounds, k0, nullptr);
                const int which = wn::radix_sort_pairs<uint64_t>(k0, cur, k1, other, nT, 0, 16 + l, sort_scratch, st);
                if (which) std::swap(cur, other);
            }
            WN_CUDA_C(cudaGetLastError());
            d_prim = cur;
            tm.mark(); // 2: order
        }
        if (kdx) {
            // tree arrays were written level by level above
        } else if (kd) {
            unsigned char* d_skip = nullptr;
            if (env_int("WN_KD_WIDE", 1)) {
                WN_CUDA_C(dalloc((void**)&d_skip, (size_t)b.nI));
            }
            wn::k_kd_tree<<<wn::grid_for(nT - 1), wn::kBuildThreads, 0, st>>>((int)nT, b.child, b.parent, b.slot, d_skip, opt.leaf_size);
            b.skip = d_skip;
            WN_CUDA_C(cudaGetLastError());
        } else if (nT >= 2) {
            wn::k_lbvh<<<wn::grid_for(nT - 1), wn::kBuildThreads, 0, st>>>(keys, (int)nT, b.child, b.parent, b.slot);
            WN_CUDA_C(cudaGetLastError());
        } else {
            const int h_child[2] = {1, -1}; // root -> leaf 0 (node id nI + 0 = 1)
            const int h_parent[2] = {-1, 0};
            WN_CUDA_C(cudaMemcpyAsync(b.child, h_child, sizeof(h_child), cudaMemcpyHostToDevice, st));
            WN_CUDA_C(cudaMemcpyAsync(b.parent, h_parent, sizeof(h_parent), cudaMemcpyHostToDevice, st));
            WN_CUDA_C(cudaStreamSynchronize(st));
        }
        tm.mark(); // 3: hierarchy
    } else {
        const int* child_src = child_in;
        if (refh) {
            // K3R: triangle boxes, then the level-synchronous top-down build (wn_refbuild_core.cuh); vertex indices are checked first
            wn::k_centroid_bounds<<<std::min(wn::grid_for(nT), 148 * 8), wn::kBuildThreads, 0, st>>>(d_v, d_tri, (int)nT, (int)nV, d_small, d_small + 6);
            int h_err = 0;
            WN_CUDA_C(cudaMemcpyAsync(&h_err, d_small + 6, sizeof(int), cudaMemcpyDeviceToHost, st));
            WN_CUDA_C(cudaStreamSynchronize(st));
            if (h_err) return cleanup(fail(WN_ERR_INVALID_ARGUMENT, "triangle references a vertex index outside [0, num_vertices)"));
            tm.mark(); // 1: vertex indices checked
            tm.mark(); // 2: (no sort in this builder)
            const size_t scan_bytes = (size_t)wn::scan_scratch_elems(nT) * 4 + 256;
            const WnRefLayout L = wn_ref_layout(nT, scan_bytes);
            char* base = nullptr;
            void* sort_scratch = nullptr;
            int* d_child_out = nullptr;
            WN_CUDA_C(dalloc((void**)&base, L.total));
            WN_CUDA_C(dalloc(&sort_scratch, (size_t)wn::sort_scratch_bytes(nT)));
            WN_CUDA_C(dalloc((void**)&d_child_out, ((size_t)nT + 2) * 16));
            WnRefState rs;
            memset(&rs, 0, sizeof(rs));
            rs.N = (int)nT;
            rs.tbox = (const float4*)(base + L.tbox);
            rs.idx = (unsigned*)(base + L.idx);
            rs.idx_alt = (unsigned*)(base + L.idx_alt);
            rs.owner = (int*)(base + L.owner);
            rs.flag = (uint32_t*)(base + L.flag);
            rs.nodes = (WnRefNode*)(base + L.nodes);
            rs.next = (WnRefNode*)(base + L.next);
            rs.tasks = (WnRefTask*)(base + L.tasks);
            rs.rows = (int*)(base + L.rows);
            rs.copen = (uint32_t*)(base + L.copen);
            rs.cnode = (uint32_t*)(base + L.cnode);
            rs.child_tmp = (int*)(base + L.child_tmp);
            rs.info_start = (int*)(base + L.info_start);
            rs.info_depth = (int*)(base + L.info_depth);
            rs.info_chain = (int*)(base + L.info_chain);
            rs.cnt_start = (uint32_t*)(base + L.cnt_start);
            rs.final_of = (int*)(base + L.final_of);
            rs.keys = (uint64_t*)(base + L.keys);
            rs.res = (int*)(base + L.res);
            rs.child_out = d_child_out;
            wn::RefCudaBackend B;
            B.st = st;
            B.scan_scratch = (uint32_t*)(base + L.scan_scratch);
            B.sort_scratch = sort_scratch;
            B.keys_alt = (uint64_t*)(base + L.keys_alt);
            B.for_each(nT, WnRefTriBoxes{d_v, d_tri, (float4*)(base + L.tbox), rs.idx, rs.owner});
            int n_nodes = 0, n_levels = 0;
            const bool ok = wn_ref_build_topology(B, rs, &n_nodes, &n_levels);
            WN_CUDA_C(cudaGetLastError());
            WN_CUDA_C(B.err);
            if (!ok) return cleanup(fail(WN_ERR_CUDA, "reference hierarchy build did not terminate"));
            num_nodes = n_nodes;
            child_src = d_child_out;
            if (env_int("WN_VERBOSE", 0)) fprintf(stderr, "[wn_b200] reference hierarchy: %d nodes, %d levels, %d host syncs\n", n_nodes, n_levels, B.syncs);
        }
        b.W = width;
        b.nI = (int)num_nodes;
        const int64_t nN = (int64_t)b.nI + b.nL;
        int* d_child_in = nullptr;
        int* d_seen = nullptr;
        WN_CUDA_C(dalloc((void**)&d_child_in, (size_t)b.nI * b.W * sizeof(int)));
        WN_CUDA_C(dalloc((void**)&d_seen, (size_t)nN * sizeof(int)));
        WN_CUDA_C(dalloc((void**)&d_prim, (size_t)nT * sizeof(unsigned)));
        WN_CUDA_C(dalloc((void**)&b.child, (size_t)b.nI * b.W * sizeof(int)));
        WN_CUDA_C(dalloc((void**)&b.parent, (size_t)nN * sizeof(int)));
        WN_CUDA_C(dalloc((void**)&b.slot, (size_t)nN));
        WN_CUDA_C(cudaMemcpyAsync(d_child_in, child_src, (size_t)b.nI * b.W * sizeof(int), cudaMemcpyDefault, st));
        WN_CUDA_C(cudaMemsetAsync(d_seen, 0, (size_t)nN * sizeof(int), st));
        WN_CUDA_C(cudaMemsetAsync(b.parent, 0xff, (size_t)nN * sizeof(int), st));
        WN_CUDA_C(cudaMemsetAsync(b.slot, 0, (size_t)nN, st));
        b.err = d_small + 6;
        // vertex index validation (the LBVH path does it in k_centroid_bounds)
        if (!refh) {
            wn::k_centroid_bounds<<<std::min(wn::grid_for(nT), 148 * 8), wn::kBuildThreads, 0, st>>>(d_v, d_tri, (int)nT, (int)nV, d_small, d_small + 6);
            tm.mark(); // 1
        }
        wn::k_iota<<<wn::grid_for(nT), wn::kBuildThreads, 0, st>>>(d_prim, (int)nT);
        if (!refh) tm.mark(); // 2
        wn::k_import_topology<<<wn::grid_for(b.nI), wn::kBuildThreads, 0, st>>>(b, d_child_in, d_seen);
        wn::k_import_check<<<wn::grid_for(nN), wn::kBuildThreads, 0, st>>>(b, d_seen);
        WN_CUDA_C(cudaGetLastError());
        int h_err = 0;
        WN_CUDA_C(cudaMemcpyAsync(&h_err, d_small + 6, sizeof(int), cudaMemcpyDeviceToHost, st));
        WN_CUDA_C(cudaStreamSynchronize(st));
        if (h_err == 3) return cleanup(fail(WN_ERR_INVALID_ARGUMENT, "triangle references a vertex index outside [0, num_vertices)"));
        if (h_err) return cleanup(fail(WN_ERR_INVALID_ARGUMENT, "malformed hierarchy topology: bad child slot, or a node/triangle not referenced exactly once"));
        tm.mark(); // 3
    }
    b.prim = d_prim;
    WN_TRY(finish_build(e, b, arena, blob_cap, tm, st)); // marks 4 (moments) and 5 (pack)

    fill_info(e);
    e->info.build_scratch_bytes = (int64_t)arena.used;
    e->info.build_ms = tm.ms(0, 5);
    e->info.build_ms_morton = tm.ms(0, 1);
    e->info.build_ms_sort = tm.ms(1, 2);
    e->info.build_ms_hierarchy = tm.ms(2, 3);
    e->info.build_ms_moments = tm.ms(3, 4);
    e->info.build_ms_pack = tm.ms(4, 5);
    // leaf entries = entries whose R2 sign bit is set: count on the host from n_entries bookkeeping is not available,
    // so derive it from sizes: for 1-triangle leaves it is nT; for collapsed leaves count on the device lazily (debug).
    e->info.num_leaf_entries = opt.leaf_size == 1 ? nT : -1;

    if (opt.keep_build_data) {
        // the kept arrays live in the build arena (or in fallback allocations): keep all of it alive with the engine
        e->kept_local = b.local;
        e->kept_child = b.child;
        e->kept_prim = d_prim;
        e->kept_nI = b.nI;
        e->kept_nL = b.nL;
        e->kept_W = b.W;
        e->kept_allocs = temps;
        temps.clear();
        if (arena.base) e->kept_allocs.push_back(arena.base);
        keep_arena = true;
    }
    *out = e;
    return cleanup(WN_OK);
#undef WN_TRY
#undef WN_CUDA_C
}

// ---- query plumbing --------------------------------------------------------------------------------------------
int env_int(const char* name, int dflt)
{
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

struct OutBufs
{
    float* d_omega = nullptr;
    uint8_t* d_inside = nullptr;
    float* h_omega = nullptr;
    uint8_t* h_inside = nullptr;
    // WN_QUERY_OUT_BITS: the kernels write one byte per query into d_inside (engine scratch), k_pack_bits folds 8 of them
    // into one byte of d_bits (the caller's device buffer, or scratch that is then copied to h_inside)
    bool bits = false;
    uint8_t* d_bits = nullptr;
};

// Resolve output residency: device pointers are used directly, host pointers get device staging.
wn_status prepare_outputs(const wn_engine* e, int64_t n, float* out_omega, uint8_t* out_inside, OutBufs& ob, bool bits = false)
{
    if (out_omega) {
        if (is_device_pointer(out_omega)) {
            ob.d_omega = out_omega;
        } else {
            WN_CUDA(e->s_out_f.reserve((size_t)n * sizeof(float)));
            ob.d_omega = (float*)e->s_out_f.p;
            ob.h_omega = out_omega;
        }
    }
    if (out_inside && bits) {
        ob.bits = true;
        WN_CUDA(e->s_out_b.reserve((size_t)n + 64));
        ob.d_inside = (uint8_t*)e->s_out_b.p;
        if (is_device_pointer(out_inside)) {
            ob.d_bits = out_inside;
        } else {
            WN_CUDA(e->s_out_bits.reserve((size_t)(n + 7) / 8 + 64));
            ob.d_bits = (uint8_t*)e->s_out_bits.p;
            ob.h_inside = out_inside;
        }
    } else if (out_inside) {
        if (is_device_pointer(out_inside)) {
            ob.d_inside = out_inside;
        } else {
            WN_CUDA(e->s_out_b.reserve((size_t)n));
            ob.d_inside = (uint8_t*)e->s_out_b.p;
            ob.h_inside = out_inside;
        }
    }
    return WN_OK;
}

// queries [first, first + count) of the batch, first a multiple of 8: bytes -> bits (LSB first: query i = bit i & 7 of byte i >> 3)
void pack_bits(const OutBufs& ob, int64_t first, int64_t count, cudaStream_t st)
{
    const int64_t nbytes = (count + 7) / 8;
    if (nbytes > 0) wn::k_pack_bits<<<(int)((nbytes + 255) / 256), 256, 0, st>>>(ob.d_inside + first, count, ob.d_bits + first / 8);
}

wn_status finish_outputs(int64_t n, const OutBufs& ob, cudaStream_t st)
{
    bool sync = false;
    if (ob.h_omega) {
        WN_CUDA(cudaMemcpyAsync(ob.h_omega, ob.d_omega, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
        sync = true;
    }
    if (ob.bits) {
        pack_bits(ob, 0, n, st);
        WN_CUDA(cudaGetLastError());
    }
    if (ob.h_inside) {
        if (ob.bits)
            WN_CUDA(cudaMemcpyAsync(ob.h_inside, ob.d_bits, (size_t)(n + 7) / 8, cudaMemcpyDeviceToHost, st));
        else
            WN_CUDA(cudaMemcpyAsync(ob.h_inside, ob.d_inside, (size_t)n, cudaMemcpyDeviceToHost, st));
        sync = true;
    }
    if (sync) WN_CUDA(cudaStreamSynchronize(st));
    return WN_OK;
}

wn_status stage_points(const wn_engine* e, const float* q_xyz, int64_t n, const float** d_q, cudaStream_t st)
{
    if (is_device_pointer(q_xyz)) {
        *d_q = q_xyz;
        return WN_OK;
    }
    WN_CUDA(e->s_in.reserve((size_t)n * 3 * sizeof(float)));
    WN_CUDA(cudaMemcpyAsync(e->s_in.p, q_xyz, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    *d_q = (const float*)e->s_in.p;
    return WN_OK;
}

template <bool GRID, bool STATS>
void launch_query(int qpl, int blocks, const wn::QueryArgs& a, cudaStream_t st)
{
    if (qpl == 2)
        wn::k_query<2, GRID, STATS><<<blocks, wn::kQueryThreads, 0, st>>>(a);
    else
        wn::k_query<1, GRID, STATS><<<blocks, wn::kQueryThreads, 0, st>>>(a);
}

int pick_qpl(int64_t n)
{
    const int forced = env_int("WN_QPL", 0);
    if (forced == 1 || forced == 2) return forced;
    return n >= (1 << 20) ? 2 : 1;
}

// The tiled path (k_tile_plan + k_tile_query) pays off when a tile's 512 queries are spatial neighbours: lattices
// always, point sets when they are Morton-sorted (by us) or declared presorted by the caller.
// 0 = generic traversal, 1 = tiled, 2 = decide by probing (sampled tile classification, see dispatch_query)
int want_tiling(const wn_engine* e, int64_t n, uint32_t flags, bool coherent)
{
    const int forced = env_int("WN_TILE", -1);
    if (forced == 0 || (flags & WN_QUERY_NO_TILING) || e->view.n_entries <= 1 || !coherent) return 0;
    if (forced == 1) return 1;
    return n >= env_int("WN_TILE_MIN", 1 << 15) ? 2 : 0;
}

float tile_kappa()
{
    const char* s = getenv("WN_KAPPA");
    const float k = s && *s ? (float)atof(s) : 6.0f;
    return k >= 1.5f ? k : 1.5f;
}

// Runs one batch of queries described by `a` (grid or points), generic or tiled, with optional executed-work counters.
// grid_layers = z1 - z0 for lattices; ignored for points.
// If `ob` has host outputs and the work is split in batches (tiled lattice), each batch's results are copied to the host
// on a second stream while the next batch computes; *copied tells the caller that nothing is left to copy.
template <bool GRID>
wn_status dispatch_query(const wn_engine* e, wn::QueryArgs& a, int64_t n, int64_t grid_layers, int tile_mode, wn_query_stats* stats,
                         cudaStream_t st, const OutBufs* ob = nullptr, bool* copied = nullptr)
{
    if (copied) *copied = false;
    bool tiled = tile_mode == 1;
    // The decision for a lattice is remembered per engine (key: lattice descriptor + beta), so that a caller that walks the
    // same lattice slab by slab (multi-GPU sharding, streaming) pays for the probe and its host synchronisation once.
    float key[8] = {0, 0, 0, 0, 0, 0, a.beta2, tile_kappa()};
    int kdims[3] = {0, 0, 0};
    if (GRID) {
        key[0] = a.g.ox, key[1] = a.g.oy, key[2] = a.g.oz, key[3] = a.g.sx, key[4] = a.g.sy, key[5] = a.g.sz;
        kdims[0] = a.g.nx, kdims[1] = a.g.ny, kdims[2] = a.g.nz;
        if (tile_mode == 2 && e->probe_cache_valid && memcmp(key, e->probe_key, sizeof(key)) == 0 && memcmp(kdims, e->probe_dims, sizeof(kdims)) == 0) {
            tiled = e->probe_cache_tiled;
            tile_mode = tiled ? 1 : 0;
        }
    }
    if (tile_mode == 2) {
        // Probe: classify ~1000 tiles spread over the batch (breadth-first part of k_tile_plan only) and estimate the share of
        // far-field evaluations the tiled path would take off every query: far / (far + direct + ~half the conditional ones).
        // Tiling pays off when that share is large (measured: cfg2 0.75 -> 1.9x faster; cfg5 0.13 -> +10 %; cfg1 0.03 -> slower).
        const int64_t total_tiles = GRID ? (int64_t)a.tiles_x * a.tiles_y * ((grid_layers + 7) / 8) : (n + wn::kTileQueries - 1) / wn::kTileQueries;
        const int stride = (int)std::max<int64_t>(1, total_tiles / 1024);
        const int blocks = (int)((total_tiles + stride - 1) / stride);
        WN_CUDA(e->s_stats.reserve(16 * sizeof(unsigned long long)));
        unsigned long long* d_probe = (unsigned long long*)e->s_stats.p + 8;
        WN_CUDA(cudaMemsetAsync(d_probe, 0, 5 * sizeof(unsigned long long), st));
        wn::QueryArgs p = a;
        p.probe_stride = stride;
        p.probe = d_probe;
        p.tile_base = 0;
        p.kappa = tile_kappa();
        p.stats = nullptr;
        wn::k_tile_plan<GRID><<<blocks, wn::kPlanThreads, 0, st>>>(p);
        unsigned long long h[5];
        WN_CUDA(cudaMemcpyAsync(h, d_probe, sizeof(h), cudaMemcpyDeviceToHost, st));
        WN_CUDA(cudaStreamSynchronize(st));
        const double far = (double)h[0], cond = (double)h[1], dir = (double)h[2];
        const double share = far / (far + dir + 0.5 * cond + 1e-9);
        const char* thr = getenv("WN_TILE_MIN_SHARE");
        // threshold re-measured at the end of round 2 (hierarchical planning, 9 plan CTAs per SM: planning got cheaper): cfg3 share
        // 0.39 -> tiled +18 %, cfg5 0.13 -> +10 %, cfg1 0.03 -> -7 %; cfg4: every sampled tile overflows -> generic
        tiled = share >= (thr && *thr ? atof(thr) : 0.10) && (double)h[4] < 0.5 * blocks;
        e->last_probe_share = (float)share;
        if (GRID) {
            memcpy(e->probe_key, key, sizeof(key));
            memcpy(e->probe_dims, kdims, sizeof(kdims));
            e->probe_cache_tiled = tiled;
            e->probe_cache_valid = true;
        }
        if (env_int("WN_VERBOSE", 0))
            fprintf(stderr, "[wn_b200] tiling probe: %d tiles, far %.0f cond %.0f direct %.0f exact %.0f fallback %.0f -> share %.3f -> %s\n", blocks,
                    far, cond, dir, (double)h[3], (double)h[4], share, tiled ? "tiled" : "generic");
    }
    if (stats) {
        WN_CUDA(e->s_stats.reserve(4 * sizeof(unsigned long long)));
        WN_CUDA(cudaMemsetAsync(e->s_stats.p, 0, 4 * sizeof(unsigned long long), st));
        a.stats = (unsigned long long*)e->s_stats.p;
    }
    const int layer_first = a.tile_z0; // lattices: first tile layer (8 z-planes) of this call; 0 unless strided
    if (!tiled) {
        const int qpl = (GRID && (a.layer_step > 1 || a.shard_q > 1)) ? 2 : pick_qpl(n);
        int64_t blocks64;
        if (GRID) {
            const int tz = (int)((grid_layers + 4 * qpl - 1) / (4 * qpl));
            blocks64 = (int64_t)a.tiles_x * a.tiles_y * tz;
        } else {
            const int per_block = wn::kQueryWarps * 32 * qpl;
            blocks64 = (n + per_block - 1) / per_block;
        }
        if (blocks64 > INT_MAX) return fail(WN_ERR_UNSUPPORTED, "batch too large for one launch; split it");
        if (stats)
            launch_query<GRID, true>(qpl, (int)blocks64, a, st);
        else
            launch_query<GRID, false>(qpl, (int)blocks64, a, st);
    } else {
        // batches of tiles so that the plan scratch (8.5 KB per tile) stays around 1 GB. Host outputs: at least two batches,
        // so that the device-to-host copy of one overlaps the computation of the next (a multi-GPU rank may hold only 32768
        // tiles); not below 8192 tiles, where launch tails cost more than the copy (measured: 16384 -> -3 %, 8192 -> -7 %).
        const bool host_out = ob && (ob->h_omega || ob->h_inside);
        int64_t units, tiles_per_unit; // grid: unit = one z layer of tiles; points: unit = one tile
        if (GRID) {
            units = (grid_layers + 7) / 8;
            tiles_per_unit = (int64_t)a.tiles_x * a.tiles_y;
        } else {
            units = (n + wn::kTileQueries - 1) / wn::kTileQueries;
            tiles_per_unit = 1;
        }
        const int64_t default_tiles = 1 << 18; // 3.2 GB of packet arena at 12 KB per tile; measured on cfg2: 2^17 -> 2^18 tiles per batch +0.7 %
        const int64_t max_tiles = std::max<int64_t>(1, env_int("WN_TILE_BATCH", (int)default_tiles));
        int64_t units_per_launch = std::max<int64_t>(1, max_tiles / tiles_per_unit);
        // hierarchical planning (lattices): blocks of 2^k x 2^k x (2^k | 1) tiles above the tiles; a batch must hold whole blocks
        const int plan_levels = GRID ? std::max(0, std::min(3, env_int("WN_PLAN_LEVELS", 2))) : 0;
        const bool zgroup = GRID && a.layer_step == 1 && a.shard_q <= 1;
        // Two lanes: consecutive batches alternate between the caller's stream and a second one (own plan scratch each), so that the
        // GPU always holds kernels of two batches: when one batch's kernel drains (launch tail, and the dependent launch behind it
        // waiting for the last CTA), the other batch's CTAs take the free SMs. Needs at least two batches of a useful size.
        const int64_t total_tiles = units * tiles_per_unit;
        const int lanes = (env_int("WN_TILE_LANES", 1) >= 2 && units >= 2 && total_tiles >= env_int("WN_TILE_LANES_MIN", 1 << 14)) ? 2 : 1;
        if (lanes == 2) units_per_launch = std::min(units_per_launch, (units + 1) / 2);
        if (plan_levels > 0 && zgroup && units_per_launch >= (1 << plan_levels)) units_per_launch &= ~(int64_t)((1 << plan_levels) - 1);
        const int64_t launch_tiles = std::min(units, units_per_launch) * tiles_per_unit;
        if (launch_tiles > INT_MAX / 2)
            return fail(WN_ERR_UNSUPPORTED, "lattice layer too large for the tiled path; pass WN_QUERY_NO_TILING or split it");
        // Variable-size packets (conditional list + gathered direct records + gathered exact triangles) are bump-allocated
        // from an arena sized for the average tile; a tile that does not fit falls back to the generic traversal.
        const int64_t arena_bytes = std::min<int64_t>(std::max<int64_t>(launch_tiles * env_int("WN_TILE_ARENA_PER_TILE", 12288), (int64_t)64 << 20),
                                                      (int64_t)4 << 30);
        for (int ln = 0; ln < lanes; ++ln) {
            wn_engine::PlanScratch& ps = e->s_plan[ln];
            WN_CUDA(ps.hdr.reserve((size_t)launch_tiles * sizeof(wn::TileHeader) + 256));
            WN_CUDA(ps.items.reserve((size_t)arena_bytes));
            WN_CUDA(ps.samples.reserve((size_t)launch_tiles * wn::kTileSampleStride * sizeof(float)));
            WN_CUDA(ps.order.reserve((size_t)launch_tiles * sizeof(int)));
        }
        a.plan_arena_bytes = arena_bytes;
        a.heavy_cond = env_int("WN_TILE_HEAVY", 192);
        a.kappa = tile_kappa();
        // Batches. Device outputs: as few as the scratch allows (every batch boundary costs a launch tail). Host outputs: the same,
        // plus a SHORT last batch (an eighth of the work): the copy of everything before it overlaps its computation and only its own
        // small copy is exposed (measured: 8 equal batches lost 4 % end to end to tails; one batch leaves the whole copy exposed).
        std::vector<int64_t> batch_units;
        {
            int64_t tail = 0;
            // (small batches, e.g. one rank's share of a lattice on 8 GPUs: a second batch boundary costs more than the ~2 MB copy)
            if (GRID && host_out && units >= 4 && units * tiles_per_unit > env_int("WN_TILE_SPLIT_MIN", 1 << 16)) {
                tail = std::max<int64_t>(1, units / std::max(2, env_int("WN_TILE_TAIL_DIV", 8)));
                if (plan_levels > 0 && zgroup) {
                    // planning blocks span 2^levels tile layers: the last batch has to start on a block boundary
                    const int64_t g = (int64_t)1 << plan_levels, start = (units - tail) / g * g;
                    tail = start > 0 ? units - start : 0;
                }
            }
            for (int64_t left = units - tail; left > 0; left -= std::min(left, units_per_launch)) batch_units.push_back(std::min(left, units_per_launch));
            // (very wide layers: the short last batch may exceed what the scratch holds -- split it like the rest)
            for (int64_t left = tail; left > 0; left -= std::min(left, units_per_launch)) batch_units.push_back(std::min(left, units_per_launch));
        }
        const bool overlap = GRID && ob && (ob->h_omega || ob->h_inside) && batch_units.size() > 1;
        if (overlap && !e->copy_stream) WN_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        // Diagnostics: WN_TRACE_FILE=<path> records {start, end, SM} of every CTA of the tiled kernels of this call (tools/cta_timeline.py)
        // (error paths: whatever returns early below still joins the second lane into `st` and frees the trace buffer)
        struct Cleanup
        {
            const wn_engine* e;
            cudaStream_t st;
            bool forked = false;
            void* trace = nullptr;
            void join()
            {
                if (forked && cudaEventRecord(e->ev_join, e->lane_stream) == cudaSuccess) cudaStreamWaitEvent(st, e->ev_join, 0);
                forked = false;
            }
            ~Cleanup()
            {
                join();
                if (trace) cudaFree(trace);
            }
        } cleanup{e, st};
        const char* trace_path = getenv("WN_TRACE_FILE");
        struct TraceLaunch { int tag, lane; int64_t offset, count; };
        std::vector<TraceLaunch> trace_launches;
        unsigned long long* d_trace = nullptr;
        int64_t trace_used = 0;
        const int64_t trace_cap = trace_path && *trace_path ? 3 * total_tiles + 65536 : 0;
        if (trace_cap > 0) {
            WN_CUDA(cudaMalloc(&d_trace, (size_t)trace_cap * 32));
            cleanup.trace = d_trace;
            WN_CUDA(cudaMemsetAsync(d_trace, 0, (size_t)trace_cap * 32, st));
        }
        auto trace_slot = [&](int tag, int lane, int64_t count) -> unsigned long long* {
            if (!d_trace || trace_used + count > trace_cap) return nullptr;
            trace_launches.push_back({tag, lane, trace_used, count});
            trace_used += count;
            return d_trace + (trace_used - count) * 4;
        };
        cudaStream_t lane_st[2] = {st, st};
        const bool forked = lanes == 2 && batch_units.size() > 1;
        if (forked) {
            if (!e->lane_stream) WN_CUDA(cudaStreamCreateWithFlags(&e->lane_stream, cudaStreamNonBlocking));
            if (!e->ev_fork) WN_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
            if (!e->ev_join) WN_CUDA(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
            WN_CUDA(cudaEventRecord(e->ev_fork, st)); // the inputs (uploaded or sorted points, counters) are ready on `st`
            WN_CUDA(cudaStreamWaitEvent(e->lane_stream, e->ev_fork, 0));
            lane_st[1] = e->lane_stream;
            cleanup.forked = true;
        }
        int64_t u0 = 0;
        for (size_t bi = 0; bi < batch_units.size(); u0 += batch_units[bi], ++bi) {
            const int64_t nunits = batch_units[bi];
            const int blocks = (int)(nunits * tiles_per_unit);
            const int ln = forked ? (int)(bi & 1) : 0;
            cudaStream_t bs = lane_st[ln];
            {
                const wn_engine::PlanScratch& ps = e->s_plan[ln];
                a.plan_hdr = (wn::TileHeader*)((char*)ps.hdr.p + 256);
                a.plan_cursor = (unsigned long long*)ps.hdr.p;
                a.plan_arena = (char*)ps.items.p;
                a.plan_samples = (float*)ps.samples.p;
                a.tile_order = (int*)ps.order.p;
            }
            if (GRID) {
                a.tile_z0 = layer_first + (int)u0 * a.layer_step;
                a.out_layer0 = (int)u0;
            } else {
                a.tile_base = u0;
            }
            a.launch_tiles = blocks;
            WN_CUDA(cudaMemsetAsync(a.plan_cursor, 0, 2 * sizeof(unsigned long long), bs)); // arena cursor + the two order counters
            a.up_hdr = nullptr;
            a.up_samples = nullptr;
            if (plan_levels > 0 && (int64_t)a.tiles_x * a.tiles_y * nunits >= 64) {
                // coarse to fine: every level reads the one above it, the tiles read level 1
                int64_t nb[4] = {0, 0, 0, 0}, total = 0;
                int lbx[4], lby[4];
                for (int k = 1; k <= plan_levels; ++k) {
                    lbx[k] = (a.tiles_x + (1 << k) - 1) >> k;
                    lby[k] = (a.tiles_y + (1 << k) - 1) >> k;
                    const int64_t lbz = zgroup ? (nunits + (1 << k) - 1) >> k : nunits;
                    nb[k] = (int64_t)lbx[k] * lby[k] * lbz;
                    total += nb[k];
                }
                const size_t hdr_bytes = align_up((size_t)total * sizeof(wn::PlanBlockHeader), 256);
                WN_CUDA(e->s_plan[ln].lvl.reserve(hdr_bytes + (size_t)total * wn::kTileSampleStride * sizeof(float)));
                wn::PlanBlockHeader* hdr_base = (wn::PlanBlockHeader*)e->s_plan[ln].lvl.p;
                float* samp_base = (float*)((char*)e->s_plan[ln].lvl.p + hdr_bytes);
                int64_t first[4] = {0, 0, 0, 0};
                for (int k = 2; k <= plan_levels; ++k) first[k] = first[k - 1] + nb[k - 1];
                for (int k = plan_levels; k >= 1; --k) {
                    wn::QueryArgs b = a;
                    b.lvl_shift = k;
                    b.lvl_zshift = zgroup ? k : 0;
                    b.lvl_bx = lbx[k];
                    b.lvl_by = lby[k];
                    b.lvl_hdr = hdr_base + first[k];
                    b.lvl_samples = samp_base + first[k] * wn::kTileSampleStride;
                    if (k < plan_levels) {
                        b.up_hdr = hdr_base + first[k + 1];
                        b.up_samples = samp_base + first[k + 1] * wn::kTileSampleStride;
                        b.up_bx = lbx[k + 1];
                        b.up_by = lby[k + 1];
                        b.up_zs = zgroup ? 1 : 0;
                    }
                    b.trace = trace_slot(10 + k, ln, nb[k]);
                    wn::k_plan_block<<<(int)nb[k], wn::kPlanThreads, 0, bs>>>(b);
                }
                a.up_hdr = hdr_base + first[1];
                a.up_samples = samp_base + first[1] * wn::kTileSampleStride;
                a.up_bx = lbx[1];
                a.up_by = lby[1];
                a.up_zs = zgroup ? 1 : 0;
            }
            a.trace = trace_slot(1, ln, blocks);
            wn::k_tile_plan<GRID><<<blocks, wn::kPlanThreads, 0, bs>>>(a);
            // a CTA walks a run of consecutive tiles, its warps taking sub-blocks dynamically (see k_tile_query)
            a.launch_tiles = blocks;
            if (ln == 0) e->last_plan_tiles = blocks;
            // A CTA takes a run of consecutive tiles of the heavy-first order and its 8 warps pull the 8 x run sub-block tasks from a
            // CTA-local counter: with one tile per CTA every warp gets exactly one task and the CTA lives as long as its slowest
            // warp. Measured on cfg2 (131072 tiles per launch): run 1 / 2 / 4 / 6 / 8 / 12 -> 7.56 / 7.81 / 7.94 / 7.96 / 7.95 / 7.92
            // G q/s; with 32768 tiles (one rank of eight): 1 / 2 / 4 / 8 -> 3.87 / 4.01 / 4.04 / 3.83 (long runs lengthen the tail).
            const int run_auto = (int)std::min<int64_t>(6, std::max<int64_t>(1, (int64_t)blocks / 8192));
            const int run_env = env_int("WN_TILE_RUN", 0);
            a.tiles_per_cta = run_env > 0 ? run_env : run_auto;
            // ... and the last WN_TILE_TAIL x (resident CTAs) x run tiles go one per CTA (CTA timeline, tools/cta_timeline.py: with
            // uniform runs the launch tail was one run of light tiles long, 5 % of a 32768-tile launch)
            if (e->query_slots == 0) {
                int sms = 148, per_sm = 5;
                WN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device));
                WN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wn::k_tile_query<GRID, false>, wn::kQueryThreads, 0));
                e->query_slots = sms * std::max(1, per_sm);
            }
            const int64_t tail_tiles =
                std::min<int64_t>(blocks, (int64_t)env_int("WN_TILE_TAIL", 2) * e->query_slots * (a.tiles_per_cta > 1 ? a.tiles_per_cta : 0));
            a.run_ctas = (int)((blocks - tail_tiles) / a.tiles_per_cta);
            const int qblocks = a.run_ctas + (blocks - a.run_ctas * a.tiles_per_cta);
            a.trace = trace_slot(2, ln, qblocks);
            if (stats)
                wn::k_tile_query<GRID, true><<<qblocks, wn::kQueryThreads, 0, bs>>>(a);
            else
                wn::k_tile_query<GRID, false><<<qblocks, wn::kQueryThreads, 0, bs>>>(a);
            if (overlap) {
                // results of z layers [8*u0, 8*(u0+nunits)) of the slab are final: ship them while the next batch runs
                const int64_t per_layer = (int64_t)a.g.nx * a.part_ny;
                const int64_t first = 8 * u0 * per_layer;
                const int64_t count = std::min<int64_t>(8 * nunits, grid_layers - 8 * u0) * per_layer;
                if (ob->bits) pack_bits(*ob, first, count, bs); // first = 8 * u0 * per_layer: byte aligned
                if (!e->ev_batch) WN_CUDA(cudaEventCreateWithFlags(&e->ev_batch, cudaEventDisableTiming));
                WN_CUDA(cudaEventRecord(e->ev_batch, bs)); // re-recording is fine: the wait below captures this record
                WN_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_batch, 0));
                if (ob->h_omega)
                    WN_CUDA(cudaMemcpyAsync(ob->h_omega + first, ob->d_omega + first, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost,
                                            e->copy_stream));
                if (ob->h_inside && ob->bits)
                    WN_CUDA(cudaMemcpyAsync(ob->h_inside + first / 8, ob->d_bits + first / 8, (size_t)(count + 7) / 8, cudaMemcpyDeviceToHost,
                                            e->copy_stream));
                else if (ob->h_inside)
                    WN_CUDA(cudaMemcpyAsync(ob->h_inside + first, ob->d_inside + first, (size_t)count, cudaMemcpyDeviceToHost, e->copy_stream));
            }
        }
        cleanup.join(); // everything after this call on `st` (bit packing, copies, the next call) sees both lanes' results
        if (d_trace) {
            // file: "WNTR", launches, then per launch {tag (1 tile plan, 2 tile query, 10+k block plan of level k), lane, count, count x 4 u64}
            a.trace = nullptr;
            std::vector<unsigned long long> h((size_t)trace_used * 4);
            WN_CUDA(cudaStreamSynchronize(st));
            WN_CUDA(cudaMemcpy(h.data(), d_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            if (FILE* f = fopen(trace_path, "wb")) {
                const int32_t nl = (int32_t)trace_launches.size();
                fwrite("WNTR", 1, 4, f);
                fwrite(&nl, sizeof(nl), 1, f);
                for (const TraceLaunch& tl : trace_launches) {
                    const int32_t hd[2] = {tl.tag, tl.lane};
                    fwrite(hd, sizeof(hd), 1, f);
                    fwrite(&tl.count, sizeof(tl.count), 1, f);
                    fwrite(h.data() + tl.offset * 4, sizeof(unsigned long long), (size_t)tl.count * 4, f);
                }
                fclose(f);
            }
        }
        if (overlap) {
            WN_CUDA(cudaGetLastError());
            WN_CUDA(cudaStreamSynchronize(e->copy_stream));
            if (copied) *copied = true;
        }
    }
    WN_CUDA(cudaGetLastError());
    if (stats) {
        unsigned long long h[4];
        WN_CUDA(cudaMemcpyAsync(h, e->s_stats.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        WN_CUDA(cudaStreamSynchronize(st));
        stats->node_tests = h[0];
        stats->far_field_evals = h[1];
        stats->exact_triangles = h[2];
        stats->lane_slots = h[3];
    }
    return WN_OK;
}

// Orders the work of consecutive calls that arrive on different streams (they share the engine's scratch buffers). Declared
// after the engine lock is taken, so it records its event before the lock is released.
struct StreamOrder
{
    const wn_engine* e;
    cudaStream_t st;
    StreamOrder(const wn_engine* e_, cudaStream_t st_) : e(e_), st(st_)
    {
        if (e->ev_last_valid && e->last_stream != st) cudaStreamWaitEvent(st, e->ev_last, 0);
    }
    ~StreamOrder()
    {
        if (!e->ev_last && cudaEventCreateWithFlags(&e->ev_last, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            e->ev_last = nullptr;
            return;
        }
        if (cudaEventRecord(e->ev_last, st) == cudaSuccess) {
            e->last_stream = st;
            e->ev_last_valid = true;
        } else {
            cudaGetLastError();
        }
    }
};

// K9: Morton order of a batch of query points (null for small batches), so that a warp's points are spatial neighbours.
// The permutation lives in the engine's sort scratch until the next call.
wn_status morton_order(const wn_engine* e, const float* d_q, int64_t n, cudaStream_t st, const unsigned** perm_out)
{
    *perm_out = nullptr;
    const int64_t sort_min = env_int("WN_SORT_MIN", 4096);
    if (n < sort_min || n > (int64_t)UINT32_MAX) return WN_OK;
    const size_t kb = align_up((size_t)n * sizeof(uint32_t), 256);
    const size_t need = 4 * kb + (size_t)wn::sort_scratch_bytes(n) + 256;
    WN_CUDA(e->s_sort.reserve(need));
    char* base = (char*)e->s_sort.p;
    uint32_t* k0 = (uint32_t*)base;
    uint32_t* k1 = (uint32_t*)(base + kb);
    unsigned* v0 = (unsigned*)(base + 2 * kb);
    unsigned* v1 = (unsigned*)(base + 3 * kb);
    int* bounds = (int*)(base + 4 * kb);
    void* scratch = base + 4 * kb + 256;
    const int init[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
    WN_CUDA(cudaMemcpyAsync(bounds, init, sizeof(init), cudaMemcpyHostToDevice, st));
    wn::k_point_bounds<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(d_q, n, bounds);
    wn::k_point_morton<<<(int)((n + 255) / 256), 256, 0, st>>>(d_q, n, bounds, k0, v0);
    const int which = wn::radix_sort_pairs<uint32_t>(k0, v0, k1, v1, n, 0, 30, scratch, st);
    WN_CUDA(cudaGetLastError());
    *perm_out = which ? v1 : v0;
    return WN_OK;
}

wn_status points_impl(const wn_engine* e, const float* q_xyz, int64_t n, float beta, uint32_t flags, float* out_omega,
                      uint8_t* out_inside, wn_query_stats* stats, void* stream)
{
    if (!e) return fail(WN_ERR_INVALID_ARGUMENT, "engine is null");
    if (n < 0 || (n > 0 && !q_xyz)) return fail(WN_ERR_INVALID_ARGUMENT, "null query buffer or negative count");
    if (!out_omega && !out_inside && !stats) return fail(WN_ERR_INVALID_ARGUMENT, "no output requested");
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n == 0) return WN_OK;
    DeviceGuard guard(e->device);
    if (!guard.ok) return fail(WN_ERR_CUDA, "cannot select CUDA device %d", e->device);
    std::lock_guard<std::mutex> lock(e->mu);
    cudaStream_t st = (cudaStream_t)stream;
    StreamOrder order(e, st);
    const float b = beta > 0.0f ? beta : e->opt.accuracy_scale;
    // Small host batches (the reference's own calling pattern is one point per call, FastWindingNumber.cpp:60-76): the kernel
    // reads the queries from, and writes the results to, pinned host memory directly (unified addressing), so a call is one
    // launch and one synchronisation instead of two staged copies around it. Same kernel, same arithmetic.
    if (n <= 256 && !stats && !(flags & WN_QUERY_OUT_BITS) && e->view.n_entries > 0 && !is_device_pointer(q_xyz) && (!out_omega || !is_device_pointer(out_omega)) &&
        (!out_inside || !is_device_pointer(out_inside))) {
        const size_t nq = (size_t)n;
        WN_CUDA(e->p_small.reserve(256 * (3 * sizeof(float) + sizeof(float) + 1)));
        float* pin_q = (float*)e->p_small.p;
        float* pin_om = pin_q + 3 * 256;
        uint8_t* pin_in = (uint8_t*)(pin_om + 256);
        memcpy(pin_q, q_xyz, nq * 3 * sizeof(float));
        wn::QueryArgs a;
        memset(&a, 0, sizeof(a));
        a.tree = e->view;
        a.beta2 = b * b;
        a.q = pin_q;
        a.n = n;
        a.q_aligned16 = 1;
        a.out_omega = out_omega ? pin_om : nullptr;
        a.out_inside = out_inside ? pin_in : nullptr;
        launch_query<false, false>(1, (int)((n + wn::kQueryWarps * 32 - 1) / (wn::kQueryWarps * 32)), a, st);
        WN_CUDA(cudaGetLastError());
        WN_CUDA(cudaStreamSynchronize(st));
        if (out_omega) memcpy(out_omega, pin_om, nq * sizeof(float));
        if (out_inside) memcpy(out_inside, pin_in, nq);
        return WN_OK;
    }
    const float* d_q = nullptr;
    wn_status s = stage_points(e, q_xyz, n, &d_q, st);
    if (s != WN_OK) return s;
    OutBufs ob;
    s = prepare_outputs(e, n, out_omega, out_inside, ob, (flags & WN_QUERY_OUT_BITS) != 0);
    if (s != WN_OK) return s;

    const unsigned* perm = nullptr;
    if (!(flags & WN_QUERY_PRESORTED) && e->view.n_entries > 0) {
        s = morton_order(e, d_q, n, st, &perm);
        if (s != WN_OK) return s;
    }

    wn::QueryArgs a;
    memset(&a, 0, sizeof(a));
    a.tree = e->view;
    a.beta2 = b * b;
    a.q = d_q;
    a.perm = perm;
    a.n = n;
    a.q_aligned16 = (((uintptr_t)d_q) & 15) == 0;
    a.out_omega = ob.d_omega;
    a.out_inside = ob.d_inside;
    const bool coherent = perm != nullptr || (flags & WN_QUERY_PRESORTED) != 0;
    s = dispatch_query<false>(e, a, n, 0, want_tiling(e, n, flags, coherent), stats, st);
    if (s != WN_OK) return s;
    return finish_outputs(n, ob, st);
}

wn_status check_grid(const float* origin, const float* spacing, const int64_t* dims, int64_t z0, int64_t z1, wn::GridDesc& g, int64_t& n)
{
    if (!origin || !spacing || !dims) return fail(WN_ERR_INVALID_ARGUMENT, "null grid description");
    if (dims[0] < 0 || dims[1] < 0 || dims[2] < 0 || dims[0] > (1 << 24) || dims[1] > (1 << 24) || dims[2] > (1 << 24))
        return fail(WN_ERR_INVALID_ARGUMENT, "grid dims must be in [0, 2^24]");
    if (z0 < 0 || z1 < z0 || z1 > dims[2]) return fail(WN_ERR_INVALID_ARGUMENT, "z slab [%lld, %lld) outside [0, %lld)", (long long)z0, (long long)z1, (long long)dims[2]);
    g.ox = origin[0];
    g.oy = origin[1];
    g.oz = origin[2];
    g.sx = spacing[0];
    g.sy = spacing[1];
    g.sz = spacing[2];
    g.nx = (int)dims[0];
    g.ny = (int)dims[1];
    g.nz = (int)dims[2];
    g.z0 = (int)z0;
    g.z1 = (int)z1;
    // dims are <= 2^24 each, so the products below cannot overflow int64; the point count itself is bounded explicitly
    if (dims[0] * dims[1] > ((int64_t)1 << 40) || dims[0] * dims[1] * (z1 - z0) > ((int64_t)1 << 40) || dims[0] * dims[1] * dims[2] > ((int64_t)1 << 42))
        return fail(WN_ERR_UNSUPPORTED, "lattice too large: at most 2^40 points per call (2^42 in the whole lattice); split it in z slabs");
    n = dims[0] * dims[1] * (z1 - z0);
    return WN_OK;
}

// layer_step == 1: the z-slab [z0, z1). layer_step > 1: the tile layers (8 z-planes each, counted from z = 0) layer_first,
// layer_first + layer_step, ... of the whole lattice, results stored compactly in that order (z0/z1 ignored).
// Diagonal sharding of a lattice over `world` ranks (wn_query_grid_sharded): the lattice is cut in Q parts along y and rank r takes,
// of every c-th tile layer (c = world / Q, starting at r mod c), the ONE part ((r - layer) / c) mod Q. Every rank sees every part and
// every height equally often, and the unit of balance is a Q-th of a layer (useful when a lattice has fewer tile layers than there are
// ranks; on cfg2, 64 layers over 8 ranks, it measured the same as whole layers within 1 %: the ranks' shares are already balanced to
// 2-3 %). Q = 4 or 2 when world and the tile rows allow it, else 1 (whole layers, the same as
// wn_query_grid_strided).
struct ShardLayout
{
    int Q = 1, c = 1;
    int64_t part_rows = 0, n_units = 0, planes = 0;
};
ShardLayout shard_layout(const int64_t* dims, int rank, int world)
{
    ShardLayout L;
    const int64_t ny = dims[1], nz = dims[2];
    for (int q : {4, 2}) {
        if (world % q == 0 && ny % (8 * q) == 0 && ny > 0) {
            L.Q = q;
            break;
        }
    }
    L.c = world / L.Q;
    L.part_rows = L.Q > 1 ? ny / L.Q : ny;
    const int64_t n_layers = (nz + 7) / 8;
    for (int64_t lz = rank % L.c; lz < n_layers; lz += L.c) {
        ++L.n_units;
        L.planes += std::min<int64_t>(8, nz - lz * 8);
    }
    return L;
}

wn_status grid_impl(const wn_engine* e, const float* origin, const float* spacing, const int64_t* dims, int64_t z0, int64_t z1, float beta,
                    uint32_t flags, float* out_omega, uint8_t* out_inside, wn_query_stats* stats, void* stream, int64_t layer_first = 0,
                    int64_t layer_step = 1, int shard_rank = 0, int shard_world = 0)
{
    if (!e) return fail(WN_ERR_INVALID_ARGUMENT, "engine is null");
    if (!out_omega && !out_inside && !stats) return fail(WN_ERR_INVALID_ARGUMENT, "no output requested");
    wn::GridDesc g;
    int64_t n = 0;
    if ((layer_step > 1 || shard_world > 1) && dims) {
        z0 = 0;
        z1 = dims[2];
    }
    wn_status s = check_grid(origin, spacing, dims, z0, z1, g, n);
    if (s != WN_OK) return s;
    int64_t local_planes = z1 - z0;
    ShardLayout sh;
    if (shard_world > 1) {
        sh = shard_layout(dims, shard_rank, shard_world);
        layer_first = shard_rank % sh.c;
        layer_step = sh.c;
        local_planes = sh.planes;
        n = dims[0] * sh.part_rows * local_planes;
    } else if (layer_step > 1) {
        if (layer_first < 0 || layer_step > (1 << 20)) return fail(WN_ERR_INVALID_ARGUMENT, "bad layer_first / layer_step");
        local_planes = 0;
        for (int64_t L = layer_first; L * 8 < dims[2]; L += layer_step) local_planes += std::min<int64_t>(8, dims[2] - L * 8);
        n = dims[0] * dims[1] * local_planes;
    }
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n == 0) return WN_OK;
    DeviceGuard guard(e->device);
    if (!guard.ok) return fail(WN_ERR_CUDA, "cannot select CUDA device %d", e->device);
    std::lock_guard<std::mutex> lock(e->mu);
    cudaStream_t st = (cudaStream_t)stream;
    StreamOrder order(e, st);
    const float b = beta > 0.0f ? beta : e->opt.accuracy_scale;
    OutBufs ob;
    s = prepare_outputs(e, n, out_omega, out_inside, ob, (flags & WN_QUERY_OUT_BITS) != 0);
    if (s != WN_OK) return s;
    wn::QueryArgs a;
    memset(&a, 0, sizeof(a));
    a.tree = e->view;
    a.beta2 = b * b;
    a.g = g;
    a.out_omega = ob.d_omega;
    a.out_inside = ob.d_inside;
    a.tiles_x = (g.nx + 7) / 8;
    a.tiles_y = (g.ny + 7) / 8;
    a.part_ny = g.ny;
    a.layer_step = (int)std::max<int64_t>(1, layer_step);
    a.tile_z0 = (layer_step > 1 || shard_world > 1) ? (int)layer_first : 0;
    a.out_layer0 = 0;
    if (shard_world > 1 && sh.Q > 1) {
        a.shard_q = sh.Q;
        a.shard_c = sh.c;
        a.shard_r = shard_rank;
        a.part_ny = (int)sh.part_rows;
        a.tiles_y = (int)(sh.part_rows / 8);
    }
    bool copied = false;
    s = dispatch_query<true>(e, a, n, local_planes, want_tiling(e, n, flags, true), stats, st, &ob, &copied);
    if (s != WN_OK) return s;
    return copied ? WN_OK : finish_outputs(n, ob, st);
}

wn_status exact_impl(const wn_engine* e, bool grid, const float* q_xyz, int64_t n, const wn::GridDesc* g, float* out_omega,
                     uint8_t* out_inside, void* stream)
{
    DeviceGuard guard(e->device);
    if (!guard.ok) return fail(WN_ERR_CUDA, "cannot select CUDA device %d", e->device);
    std::lock_guard<std::mutex> lock(e->mu);
    cudaStream_t st = (cudaStream_t)stream;
    StreamOrder order(e, st);
    const float* d_q = nullptr;
    wn_status s = WN_OK;
    if (!grid) {
        s = stage_points(e, q_xyz, n, &d_q, st);
        if (s != WN_OK) return s;
    }
    OutBufs ob;
    s = prepare_outputs(e, n, out_omega, out_inside, ob);
    if (s != WN_OK) return s;
    wn::ExactArgs a;
    memset(&a, 0, sizeof(a));
    a.tris = e->view.tri;
    a.nT = e->view.n_tris;
    a.q = d_q;
    a.n = n;
    if (g) a.g = *g;
    a.out_omega = ob.d_omega;
    a.out_inside = ob.d_inside;
    const bool small = n < 2048;
    const int64_t qblocks = small ? (n + 7) / 8 : (n + 255) / 256;
    const int ntiles = std::max(1, (a.nT + wn::kExactTile - 1) / wn::kExactTile);
    // enough blocks to fill 148 SMs a few times over; chunks are whole tiles
    const int64_t target_blocks = 148 * 16;
    int nchunks = (int)std::min<int64_t>(ntiles, std::max<int64_t>(1, (target_blocks + qblocks - 1) / qblocks));
    nchunks = std::min(nchunks, 65535);
    const int tiles_per_chunk = (ntiles + nchunks - 1) / nchunks;
    nchunks = (ntiles + tiles_per_chunk - 1) / tiles_per_chunk;
    a.tris_per_chunk = tiles_per_chunk * wn::kExactTile;
    a.nchunks = nchunks;
    if (nchunks > 1) {
        WN_CUDA(e->s_partial.reserve((size_t)nchunks * n * sizeof(float)));
        a.partial = (float*)e->s_partial.p;
    }
    if (qblocks > INT_MAX) return fail(WN_ERR_UNSUPPORTED, "too many queries for one exact launch");
    dim3 gridDim((unsigned)qblocks, (unsigned)nchunks);
    if (small) {
        if (grid)
            wn::k_exact_small<true><<<gridDim, 256, 0, st>>>(a);
        else
            wn::k_exact_small<false><<<gridDim, 256, 0, st>>>(a);
    } else {
        if (grid)
            wn::k_exact<true><<<gridDim, 256, 0, st>>>(a);
        else
            wn::k_exact<false><<<gridDim, 256, 0, st>>>(a);
    }
    if (nchunks > 1) wn::k_exact_reduce<<<(int)((n + 255) / 256), 256, 0, st>>>(a.partial, nchunks, n, ob.d_omega, ob.d_inside);
    WN_CUDA(cudaGetLastError());
    return finish_outputs(n, ob, st);
}

template <typename K>
wn_status debug_sort(K* keys, uint32_t* values, int64_t n, int32_t begin_bit, int32_t end_bit)
{
    if (n < 0 || (n > 0 && (!keys || !values))) return fail(WN_ERR_INVALID_ARGUMENT, "null buffers");
    if (begin_bit < 0 || end_bit > (int)sizeof(K) * 8 || begin_bit >= end_bit) return fail(WN_ERR_INVALID_ARGUMENT, "bad bit range");
    if (n == 0) return WN_OK;
    int dev = 0;
    wn_status s = resolve_device(nullptr, &dev);
    if (s != WN_OK) return s;
    K *k0 = nullptr, *k1 = nullptr;
    uint32_t *v0 = nullptr, *v1 = nullptr;
    void* scratch = nullptr;
    cudaError_t ce = cudaMalloc((void**)&k0, n * sizeof(K));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&k1, n * sizeof(K));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&v0, n * sizeof(uint32_t));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&v1, n * sizeof(uint32_t));
    if (ce == cudaSuccess) ce = cudaMalloc(&scratch, (size_t)wn::sort_scratch_bytes(n));
    if (ce == cudaSuccess) ce = cudaMemcpy(k0, keys, n * sizeof(K), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(v0, values, n * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) {
        const int which = wn::radix_sort_pairs<K>(k0, v0, k1, v1, n, begin_bit, end_bit, scratch, nullptr);
        ce = cudaGetLastError();
        if (ce == cudaSuccess) ce = cudaMemcpy(keys, which ? k1 : k0, n * sizeof(K), cudaMemcpyDeviceToHost);
        if (ce == cudaSuccess) ce = cudaMemcpy(values, which ? v1 : v0, n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    }
    cudaFree(k0);
    cudaFree(k1);
    cudaFree(v0);
    cudaFree(v1);
    cudaFree(scratch);
    WN_CUDA(ce);
    return WN_OK;
}

} // namespace

// =================================================================================================================
extern "C" {

const char* wn_last_error(void)
{
    return g_last_error.c_str();
}

const char* wn_version(void)
{
    return "lagrange_b200 winding 0.1 (sm_100a)";
}

wn_status wn_options_init(wn_options* opt)
{
    if (!opt) return fail(WN_ERR_INVALID_ARGUMENT, "opt is null");
    memset(opt, 0, sizeof(*opt));
    opt->struct_size = sizeof(wn_options);
    opt->device = -1;
    opt->accuracy_scale = 2.0f;
    opt->order = 2;
    opt->leaf_size = 1;
    opt->morton_bits = 63;
    opt->radius_mode = WN_RADIUS_BOX_CORNER;
    opt->approximate_single_triangles = 0;
    opt->keep_build_data = 0;
    opt->hierarchy = WN_HIERARCHY_REFERENCE; // the drop-in default: the reference builder's own tree, results match the reference algorithm
    return WN_OK;
}

wn_status wn_create(const float* v_xyz, int64_t num_vertices, const int32_t* tri, int64_t num_triangles, const wn_options* opt, wn_engine** out)
{
    return create_impl(v_xyz, num_vertices, tri, num_triangles, nullptr, 0, 0, opt, out);
}

wn_status wn_create_from_topology(const float* v_xyz, int64_t num_vertices, const int32_t* tri, int64_t num_triangles, const int32_t* child,
                                  int64_t num_nodes, int32_t width, const wn_options* opt, wn_engine** out)
{
    if (!child) return fail(WN_ERR_INVALID_ARGUMENT, "child table is null");
    return create_impl(v_xyz, num_vertices, tri, num_triangles, child, num_nodes, width, opt, out);
}

wn_status wn_destroy(wn_engine* e)
{
    if (!e) return WN_OK;
    {
        DeviceGuard guard(e->device);
        if (e->blob) cudaFree(e->blob);
        for (void* p : e->kept_allocs) cudaFree(p);
        e->s_in.release();
        e->s_out_f.release();
        e->s_out_b.release();
        e->s_out_bits.release();
        e->s_sort.release();
        e->s_stats.release();
        e->s_partial.release();
        e->s_plan[0].release();
        e->s_plan[1].release();
        e->s_sdf_inside.release();
        e->s_sdf_dense.release();
        e->p_small.release();
        if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
        if (e->lane_stream) cudaStreamDestroy(e->lane_stream);
        if (e->ev_fork) cudaEventDestroy(e->ev_fork);
        if (e->ev_join) cudaEventDestroy(e->ev_join);
        if (e->ev_last) cudaEventDestroy(e->ev_last);
        if (e->ev_batch) cudaEventDestroy(e->ev_batch);
    }
    delete e;
    return WN_OK;
}

wn_status wn_get_info(const wn_engine* e, wn_info* info)
{
    if (!e || !info) return fail(WN_ERR_INVALID_ARGUMENT, "null argument");
    *info = e->info;
    return WN_OK;
}

wn_status wn_solid_angle(const wn_engine* e, const float* q_xyz, int64_t n, float beta, uint32_t flags, float* out_omega, void* stream)
{
    if (!out_omega && n > 0) return fail(WN_ERR_INVALID_ARGUMENT, "out_omega is null");
    return points_impl(e, q_xyz, n, beta, flags, out_omega, nullptr, nullptr, stream);
}

wn_status wn_is_inside(const wn_engine* e, const float* q_xyz, int64_t n, float beta, uint32_t flags, uint8_t* out_inside, void* stream)
{
    if (!out_inside && n > 0) return fail(WN_ERR_INVALID_ARGUMENT, "out_inside is null");
    return points_impl(e, q_xyz, n, beta, flags, nullptr, out_inside, nullptr, stream);
}

wn_status wn_query_grid(const wn_engine* e, const float origin[3], const float spacing[3], const int64_t dims[3], int64_t z_begin,
                        int64_t z_end, float beta, uint32_t flags, float* out_omega, uint8_t* out_inside, void* stream)
{
    return grid_impl(e, origin, spacing, dims, z_begin, z_end, beta, flags, out_omega, out_inside, nullptr, stream);
}

wn_status wn_query_grid_strided(const wn_engine* e, const float origin[3], const float spacing[3], const int64_t dims[3], int64_t layer_first,
                                int64_t layer_step, float beta, uint32_t flags, float* out_omega, uint8_t* out_inside, void* stream)
{
    if (layer_step < 1) return fail(WN_ERR_INVALID_ARGUMENT, "layer_step must be >= 1");
    if (layer_step == 1) return grid_impl(e, origin, spacing, dims, layer_first * 8, dims ? dims[2] : 0, beta, flags, out_omega, out_inside, nullptr, stream);
    return grid_impl(e, origin, spacing, dims, 0, 0, beta, flags, out_omega, out_inside, nullptr, stream, layer_first, layer_step);
}

wn_status wn_query_grid_sharded(const wn_engine* e, const float origin[3], const float spacing[3], const int64_t dims[3], int32_t rank, int32_t world,
                                float beta, uint32_t flags, float* out_omega, uint8_t* out_inside, void* stream)
{
    if (world < 1 || rank < 0 || rank >= world || world > (1 << 16)) return fail(WN_ERR_INVALID_ARGUMENT, "rank / world out of range");
    if (world == 1) return grid_impl(e, origin, spacing, dims, 0, dims ? dims[2] : 0, beta, flags, out_omega, out_inside, nullptr, stream);
    return grid_impl(e, origin, spacing, dims, 0, 0, beta, flags, out_omega, out_inside, nullptr, stream, 0, 1, rank, world);
}

wn_status wn_grid_shard_layout(const int64_t dims[3], int32_t rank, int32_t world, int32_t* parts_y, int32_t* layer_step, int64_t* part_rows,
                               int64_t* n_units, int64_t* n_points)
{
    if (!dims || world < 1 || rank < 0 || rank >= world || dims[0] < 0 || dims[1] < 0 || dims[2] < 0)
        return fail(WN_ERR_INVALID_ARGUMENT, "bad lattice or rank / world");
    const ShardLayout L = shard_layout(dims, rank, world);
    if (parts_y) *parts_y = L.Q;
    if (layer_step) *layer_step = L.c;
    if (part_rows) *part_rows = L.part_rows;
    if (n_units) *n_units = L.n_units;
    if (n_points) *n_points = dims[0] * L.part_rows * L.planes;
    return WN_OK;
}

wn_status wn_query_stats_points(const wn_engine* e, const float* q_xyz, int64_t n, float beta, uint32_t flags, wn_query_stats* stats, void* stream)
{
    if (!stats) return fail(WN_ERR_INVALID_ARGUMENT, "stats is null");
    return points_impl(e, q_xyz, n, beta, flags, nullptr, nullptr, stats, stream);
}

wn_status wn_query_stats_grid(const wn_engine* e, const float origin[3], const float spacing[3], const int64_t dims[3], int64_t z_begin,
                              int64_t z_end, float beta, uint32_t flags, wn_query_stats* stats, void* stream)
{
    if (!stats) return fail(WN_ERR_INVALID_ARGUMENT, "stats is null");
    return grid_impl(e, origin, spacing, dims, z_begin, z_end, beta, flags, nullptr, nullptr, stats, stream);
}

wn_status wn_exact(const wn_engine* e, const float* q_xyz, int64_t n, float* out_omega, uint8_t* out_inside, void* stream)
{
    if (!e) return fail(WN_ERR_INVALID_ARGUMENT, "engine is null");
    if (n < 0 || (n > 0 && !q_xyz)) return fail(WN_ERR_INVALID_ARGUMENT, "null query buffer or negative count");
    if (!out_omega && !out_inside) return fail(WN_ERR_INVALID_ARGUMENT, "no output requested");
    if (n == 0) return WN_OK;
    return exact_impl(e, false, q_xyz, n, nullptr, out_omega, out_inside, stream);
}

wn_status wn_exact_grid(const wn_engine* e, const float origin[3], const float spacing[3], const int64_t dims[3], int64_t z_begin, int64_t z_end,
                        float* out_omega, uint8_t* out_inside, void* stream)
{
    if (!e) return fail(WN_ERR_INVALID_ARGUMENT, "engine is null");
    if (!out_omega && !out_inside) return fail(WN_ERR_INVALID_ARGUMENT, "no output requested");
    wn::GridDesc g;
    int64_t n = 0;
    wn_status s = check_grid(origin, spacing, dims, z_begin, z_end, g, n);
    if (s != WN_OK) return s;
    if (n == 0) return WN_OK;
    return exact_impl(e, true, nullptr, n, &g, out_omega, out_inside, stream);
}

wn_status wn_tree_packed_size(const wn_engine* e, int64_t* nbytes)
{
    if (!e || !nbytes) return fail(WN_ERR_INVALID_ARGUMENT, "null argument");
    *nbytes = e->hdr.total_bytes;
    return WN_OK;
}

wn_status wn_tree_pack(const wn_engine* e, void* dst, int64_t nbytes, void* stream)
{
    if (!e || !dst) return fail(WN_ERR_INVALID_ARGUMENT, "null argument");
    if (nbytes < e->hdr.total_bytes) return fail(WN_ERR_INVALID_ARGUMENT, "destination too small: need %lld bytes", (long long)e->hdr.total_bytes);
    DeviceGuard guard(e->device);
    if (!guard.ok) return fail(WN_ERR_CUDA, "cannot select CUDA device %d", e->device);
    cudaStream_t st = (cudaStream_t)stream;
    WN_CUDA(cudaMemcpyAsync(dst, e->blob, (size_t)e->hdr.total_bytes, cudaMemcpyDefault, st));
    if (!is_device_pointer(dst)) WN_CUDA(cudaStreamSynchronize(st));
    return WN_OK;
}

wn_status wn_create_from_packed(const void* src, int64_t nbytes, const wn_options* opt_in, wn_engine** out)
{
    if (!out) return fail(WN_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    if (!src || nbytes < (int64_t)sizeof(PackedHeader)) return fail(WN_ERR_INVALID_ARGUMENT, "packed tree buffer missing or too small");
    // accuracy_scale <= 0 in the options means "keep the value stored with the tree" (a replica must answer beta <= 0 queries
    // like the engine it was packed from); everything else in the options is validated as usual.
    wn_options opt_copy;
    bool keep_beta = true;
    if (opt_in) {
        if (opt_in->struct_size != sizeof(wn_options)) return fail(WN_ERR_INVALID_ARGUMENT, "wn_options.struct_size mismatch (call wn_options_init)");
        opt_copy = *opt_in;
        keep_beta = !(opt_copy.accuracy_scale > 0.0f);
        if (keep_beta) opt_copy.accuracy_scale = 2.0f; // placeholder for validation; replaced by the header's value below
    }
    wn_options opt;
    wn_status s = validate_options(opt_in ? &opt_copy : nullptr, &opt, false);
    if (s != WN_OK) return s;
    int dev = 0;
    s = resolve_device(&opt, &dev);
    if (s != WN_OK) return s;
    DeviceGuard guard(dev);
    if (!guard.ok) return fail(WN_ERR_CUDA, "cannot select CUDA device %d", dev);
    PackedHeader h;
    WN_CUDA(cudaMemcpy(&h, src, sizeof(h), cudaMemcpyDefault));
    if (h.magic != kMagic) return fail(WN_ERR_INVALID_ARGUMENT, "not a packed winding tree (bad magic)");
    if (h.n_entries < 0 || h.n_entries > INT_MAX / 8 || h.n_tris < 0 || h.n_tris >= WN_MAX_TRIANGLES)
        return fail(WN_ERR_INVALID_ARGUMENT, "packed tree header is inconsistent (entry / triangle counts out of range)");
    if (h.total_bytes > nbytes) return fail(WN_ERR_INVALID_ARGUMENT, "packed tree truncated: header says %lld bytes, got %lld", (long long)h.total_bytes, (long long)nbytes);
    const PackedHeader ref = make_header(h.n_entries, h.n_tris);
    if (ref.total_bytes != h.total_bytes || ref.off_hot != h.off_hot || ref.off_cold != h.off_cold || ref.off_kids != h.off_kids ||
        ref.off_tris != h.off_tris || ref.off_tri_order != h.off_tri_order)
        return fail(WN_ERR_INVALID_ARGUMENT, "packed tree header is inconsistent");
    if (h.order < 0 || h.order > 2 || !(h.accuracy_scale > 0.0f) || (h.n_entries == 0) != (h.n_tris == 0))
        return fail(WN_ERR_INVALID_ARGUMENT, "packed tree header is inconsistent (order / accuracy scale / empty tree)");
    wn_engine* e = new (std::nothrow) wn_engine;
    if (!e) return fail(WN_ERR_OUT_OF_MEMORY, "host allocation failed");
    e->device = dev;
    e->opt = opt;
    e->opt.accuracy_scale = keep_beta ? h.accuracy_scale : opt.accuracy_scale;
    e->opt.order = h.order;
    memset(&e->info, 0, sizeof(e->info));
    e->hdr = h;
    cudaError_t ce = cudaMalloc((void**)&e->blob, (size_t)h.total_bytes);
    if (ce == cudaSuccess) ce = cudaMemcpy(e->blob, src, (size_t)h.total_bytes, cudaMemcpyDefault);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        wn_destroy(e);
        return fail(ce == cudaErrorMemoryAllocation ? WN_ERR_OUT_OF_MEMORY : WN_ERR_CUDA, "adopting packed tree failed: %s", cudaGetErrorString(ce));
    }
    set_view(e);
    // The records are about to be dereferenced by the traversal: check every link, child index and leaf range once (a
    // corrupted or mismatched broadcast must be an error, not an out-of-bounds read). One pass over the hot section.
    if (h.n_entries > 0) {
        int* d_err = nullptr;
        ce = cudaMalloc((void**)&d_err, sizeof(int));
        if (ce == cudaSuccess) ce = cudaMemset(d_err, 0, sizeof(int));
        int h_err = 0;
        if (ce == cudaSuccess) {
            wn::k_validate_packed<<<wn::grid_for(h.n_entries), wn::kBuildThreads>>>(e->view, d_err);
            ce = cudaMemcpy(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost);
        }
        if (d_err) cudaFree(d_err);
        if (ce != cudaSuccess || h_err) {
            cudaGetLastError();
            wn_destroy(e);
            if (ce != cudaSuccess) return fail(WN_ERR_CUDA, "validating packed tree failed: %s", cudaGetErrorString(ce));
            return fail(WN_ERR_INVALID_ARGUMENT, "packed tree is corrupt (bad skip link, child index or leaf range)");
        }
    }
    e->hdr.accuracy_scale = e->opt.accuracy_scale;
    fill_info(e);
    e->info.num_leaf_entries = -1;
    *out = e;
    return WN_OK;
}

wn_status wn_debug_node_moments(const wn_engine* e, int64_t first_node, int64_t count, float* out_23)
{
    if (!e || !out_23) return fail(WN_ERR_INVALID_ARGUMENT, "null argument");
    if (!e->kept_local) return fail(WN_ERR_INVALID_ARGUMENT, "engine was built without keep_build_data");
    const int64_t nN = (int64_t)e->kept_nI + e->kept_nL;
    if (first_node < 0 || count < 0 || first_node + count > nN) return fail(WN_ERR_INVALID_ARGUMENT, "node range outside [0, %lld)", (long long)nN);
    if (count == 0) return WN_OK;
    DeviceGuard guard(e->device);
    float* d_out = nullptr;
    WN_CUDA(cudaMalloc((void**)&d_out, (size_t)count * 23 * sizeof(float)));
    WnBuild b;
    memset(&b, 0, sizeof(b));
    b.local = e->kept_local;
    wn::k_ref23<<<wn::grid_for(count), wn::kBuildThreads>>>(b, first_node, count, d_out);
    cudaError_t ce = cudaMemcpy(out_23, d_out, (size_t)count * 23 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_out);
    WN_CUDA(ce);
    return WN_OK;
}

wn_status wn_debug_topology(const wn_engine* e, int32_t* child, int64_t capacity_nodes, int64_t* num_internal)
{
    if (!e || !num_internal) return fail(WN_ERR_INVALID_ARGUMENT, "null argument");
    if (!e->kept_child) return fail(WN_ERR_INVALID_ARGUMENT, "engine was built without keep_build_data");
    *num_internal = e->kept_nI;
    if (!child) return WN_OK;
    if (capacity_nodes < e->kept_nI) return fail(WN_ERR_INVALID_ARGUMENT, "child buffer too small");
    DeviceGuard guard(e->device);
    const size_t n = (size_t)e->kept_nI * e->kept_W;
    std::vector<int> h(n);
    std::vector<unsigned> prim((size_t)e->kept_nL);
    WN_CUDA(cudaMemcpy(h.data(), e->kept_child, n * sizeof(int), cudaMemcpyDeviceToHost));
    WN_CUDA(cudaMemcpy(prim.data(), e->kept_prim, prim.size() * sizeof(unsigned), cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < n; ++k) {
        const int c = h[k];
        child[k] = c < 0 ? WN_CHILD_EMPTY : (c >= e->kept_nI ? wn_enc_tri((int)prim[(size_t)(c - e->kept_nI)]) : c);
    }
    return WN_OK;
}

wn_status wn_sdf_grid(const wn_engine* e, const float* origin, const float* spacing, const int64_t* dims, float band, float beta,
                      uint32_t flags, float* out_sdf, int64_t* num_active, void* stream)
{
    if (!e) return fail(WN_ERR_INVALID_ARGUMENT, "engine is null");
    if (!out_sdf) return fail(WN_ERR_INVALID_ARGUMENT, "no output requested");
    if (!(band > 0.0f) || !(band <= 3.0e38f)) return fail(WN_ERR_INVALID_ARGUMENT, "band must be positive and finite");
    wn::GridDesc g;
    int64_t n = 0;
    wn_status s = check_grid(origin, spacing, dims, 0, dims ? dims[2] : 0, g, n);
    if (s != WN_OK) return s;
    if (num_active) *num_active = 0;
    if (n == 0) return WN_OK;
    const int64_t tiles = (int64_t)((g.nx + 7) / 8) * ((g.ny + 7) / 8) * ((g.nz + 7) / 8);
    if (tiles > INT_MAX) return fail(WN_ERR_UNSUPPORTED, "lattice too large for one launch; split it");
    DeviceGuard guard(e->device);
    if (!guard.ok) return fail(WN_ERR_CUDA, "cannot select CUDA device %d", e->device);
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> sdf_lock(e->sdf_mu);
    uint8_t* d_inside = nullptr;
    if (!(flags & WN_SDF_UNSIGNED)) {
        {
            std::lock_guard<std::mutex> lock(e->mu);
            WN_CUDA(e->s_sdf_inside.reserve((size_t)n));
            d_inside = (uint8_t*)e->s_sdf_inside.p;
        }
        // the sign: the reference's interior test, FastWindingNumber::is_inside at every voxel centre
        s = grid_impl(e, origin, spacing, dims, 0, dims[2], beta, flags & WN_QUERY_NO_TILING, nullptr, d_inside, nullptr, stream);
        if (s != WN_OK) return s;
    }
    std::lock_guard<std::mutex> lock(e->mu);
    StreamOrder order(e, st);
    OutBufs ob;
    s = prepare_outputs(e, n, out_sdf, nullptr, ob);
    if (s != WN_OK) return s;
    WN_CUDA(e->s_stats.reserve(16 * sizeof(unsigned long long)));
    unsigned long long* d_active = (unsigned long long*)e->s_stats.p;
    WN_CUDA(cudaMemsetAsync(d_active, 0, sizeof(unsigned long long), st));
    wn::SdfArgs a;
    memset(&a, 0, sizeof(a));
    a.tree = e->view;
    a.g = g;
    a.tiles_x = (g.nx + 7) / 8;
    a.tiles_y = (g.ny + 7) / 8;
    a.band = band;
    a.inside = d_inside;
    a.out = ob.d_omega;
    a.active = d_active;
    wn::k_sdf_grid<<<(int)tiles, wn::kQueryThreads, 0, st>>>(a);
    WN_CUDA(cudaGetLastError());
    if (num_active) {
        unsigned long long h = 0;
        WN_CUDA(cudaMemcpyAsync(&h, d_active, sizeof(h), cudaMemcpyDeviceToHost, st));
        WN_CUDA(cudaStreamSynchronize(st));
        *num_active = (int64_t)h;
    }
    return finish_outputs(n, ob, st);
}

wn_status wn_closest_point(const wn_engine* e, const float* q_xyz, int64_t n, float max_distance, uint32_t flags, float* out_sqdist,
                           int32_t* out_triangle, float* out_xyz, void* stream)
{
    if (!e) return fail(WN_ERR_INVALID_ARGUMENT, "engine is null");
    if (n < 0 || (n > 0 && !q_xyz)) return fail(WN_ERR_INVALID_ARGUMENT, "null query buffer or negative count");
    if (!out_sqdist && !out_triangle && !out_xyz) return fail(WN_ERR_INVALID_ARGUMENT, "no output requested");
    if (n == 0) return WN_OK;
    if (n > ((int64_t)1 << 36)) return fail(WN_ERR_UNSUPPORTED, "too many points for one call; split it");
    DeviceGuard guard(e->device);
    if (!guard.ok) return fail(WN_ERR_CUDA, "cannot select CUDA device %d", e->device);
    std::lock_guard<std::mutex> lock(e->mu);
    cudaStream_t st = (cudaStream_t)stream;
    StreamOrder order(e, st);
    const float* d_q = nullptr;
    wn_status s = stage_points(e, q_xyz, n, &d_q, st);
    if (s != WN_OK) return s;
    // outputs: device pointers in place, host pointers through one scratch block [sqdist n | tri n | xyz 3n]
    const bool h_sq = out_sqdist && !is_device_pointer(out_sqdist), h_tri = out_triangle && !is_device_pointer(out_triangle),
               h_xyz = out_xyz && !is_device_pointer(out_xyz);
    float* d_sq = out_sqdist;
    int* d_tri = out_triangle;
    float* d_xyz = out_xyz;
    if (h_sq || h_tri || h_xyz) {
        WN_CUDA(e->s_out_f.reserve((size_t)n * 5 * sizeof(float)));
        float* base = (float*)e->s_out_f.p;
        if (h_sq) d_sq = base;
        if (h_tri) d_tri = (int*)(base + n);
        if (h_xyz) d_xyz = base + 2 * n;
    }
    const unsigned* perm = nullptr;
    if (!(flags & WN_QUERY_PRESORTED) && e->view.n_entries > 0) {
        s = morton_order(e, d_q, n, st, &perm);
        if (s != WN_OK) return s;
    }
    wn::ClosestArgs a;
    memset(&a, 0, sizeof(a));
    a.tree = e->view;
    a.tri_order = (const unsigned*)(e->blob + e->hdr.off_tri_order);
    a.q = d_q;
    a.perm = perm;
    a.n = n;
    a.max_dist = max_distance;
    a.out_sqdist = d_sq;
    a.out_tri = d_tri;
    a.out_xyz = d_xyz;
    wn::k_closest_point<<<(int)((n + wn::kQueryThreads - 1) / wn::kQueryThreads), wn::kQueryThreads, 0, st>>>(a);
    WN_CUDA(cudaGetLastError());
    if (h_sq) WN_CUDA(cudaMemcpyAsync(out_sqdist, d_sq, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (h_tri) WN_CUDA(cudaMemcpyAsync(out_triangle, d_tri, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (h_xyz) WN_CUDA(cudaMemcpyAsync(out_xyz, d_xyz, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (h_sq || h_tri || h_xyz) WN_CUDA(cudaStreamSynchronize(st));
    return WN_OK;
}

wn_status wn_sdf_grid_sparse(const wn_engine* e, const float* origin, const float* spacing, const int64_t* dims, float band, float beta,
                             uint32_t flags, int64_t capacity, int64_t* out_index, float* out_value, uint8_t* out_inside_bits, int64_t* num_active,
                             void* stream)
{
    if (!e) return fail(WN_ERR_INVALID_ARGUMENT, "engine is null");
    if (!num_active) return fail(WN_ERR_INVALID_ARGUMENT, "num_active is null");
    if (capacity < 0 || (capacity > 0 && (!out_index || !out_value))) return fail(WN_ERR_INVALID_ARGUMENT, "null output buffers");
    wn::GridDesc g;
    int64_t n = 0;
    wn_status s = check_grid(origin, spacing, dims, 0, dims ? dims[2] : 0, g, n);
    if (s != WN_OK) return s;
    *num_active = 0;
    if (n == 0) return WN_OK;
    DeviceGuard guard(e->device);
    if (!guard.ok) return fail(WN_ERR_CUDA, "cannot select CUDA device %d", e->device);
    cudaStream_t st = (cudaStream_t)stream;
    // dense block in device scratch (never leaves the GPU), then compaction of the band
    float* d_sdf = nullptr;
    {
        std::lock_guard<std::mutex> lock(e->mu);
        WN_CUDA(e->s_sdf_dense.reserve((size_t)n * sizeof(float)));
        d_sdf = (float*)e->s_sdf_dense.p;
    }
    s = wn_sdf_grid(e, origin, spacing, dims, band, beta, flags, d_sdf, nullptr, stream);
    if (s != WN_OK) return s;
    std::lock_guard<std::mutex> sdf_lock(e->sdf_mu);
    std::lock_guard<std::mutex> lock(e->mu);
    StreamOrder order(e, st);
    const int64_t per_block = (int64_t)wn::kQueryThreads * wn::kCompactItems;
    const int64_t blocks = (n + per_block - 1) / per_block;
    if (blocks > INT_MAX) return fail(WN_ERR_UNSUPPORTED, "lattice too large for one launch; split it");
    WN_CUDA(e->s_sort.reserve((size_t)(blocks + 1) * 4 + (size_t)wn::scan_scratch_elems(blocks + 1) * 4 + 512));
    uint32_t* d_counts = (uint32_t*)e->s_sort.p;
    uint32_t* d_scan = (uint32_t*)((char*)e->s_sort.p + align_up((size_t)(blocks + 1) * 4, 256));
    WN_CUDA(cudaMemsetAsync(d_counts + blocks, 0, 4, st));
    wn::k_band_count<<<(int)blocks, wn::kQueryThreads, 0, st>>>(d_sdf, n, band, d_counts);
    wn::exclusive_scan_u32(d_counts, blocks + 1, d_scan, st);
    uint32_t h_total = 0;
    WN_CUDA(cudaMemcpyAsync(&h_total, d_counts + blocks, 4, cudaMemcpyDeviceToHost, st));
    WN_CUDA(cudaStreamSynchronize(st));
    *num_active = (int64_t)h_total;
    const int64_t m = std::min<int64_t>(capacity, (int64_t)h_total);
    if (m > 0) {
        const bool h_idx = !is_device_pointer(out_index), h_val = !is_device_pointer(out_value);
        int64_t* d_idx = out_index;
        float* d_val = out_value;
        if (h_idx || h_val) {
            WN_CUDA(e->s_out_f.reserve((size_t)m * 12 + 256));
            if (h_idx) d_idx = (int64_t*)e->s_out_f.p;
            if (h_val) d_val = (float*)((char*)e->s_out_f.p + align_up((size_t)m * 8, 256));
        }
        wn::k_band_scatter<<<(int)blocks, wn::kQueryThreads, 0, st>>>(d_sdf, n, band, d_counts, m, d_idx, d_val);
        WN_CUDA(cudaGetLastError());
        if (h_idx) WN_CUDA(cudaMemcpyAsync(out_index, d_idx, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
        if (h_val) WN_CUDA(cudaMemcpyAsync(out_value, d_val, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    }
    if (out_inside_bits) {
        // sign of every voxel, 1 bit each (the inactive interior is not in the band list): taken from the dense block
        const bool h_bits = !is_device_pointer(out_inside_bits);
        uint8_t* d_bits = out_inside_bits;
        if (h_bits) {
            WN_CUDA(e->s_out_bits.reserve((size_t)(n + 7) / 8 + 64));
            d_bits = (uint8_t*)e->s_out_bits.p;
        }
        wn::k_sign_bits<<<(int)(((n + 7) / 8 + 255) / 256), 256, 0, st>>>(d_sdf, n, d_bits);
        WN_CUDA(cudaGetLastError());
        if (h_bits) WN_CUDA(cudaMemcpyAsync(out_inside_bits, d_bits, (size_t)(n + 7) / 8, cudaMemcpyDeviceToHost, st));
    }
    WN_CUDA(cudaStreamSynchronize(st));
    return WN_OK;
}

wn_status wn_replicate(const wn_engine* src, const int32_t* devices, int32_t n_devices, wn_engine** out)
{
    if (!src || !devices || !out || n_devices < 0) return fail(WN_ERR_INVALID_ARGUMENT, "null argument");
    for (int i = 0; i < n_devices; ++i) out[i] = nullptr;
    int count = 0;
    WN_CUDA(cudaGetDeviceCount(&count));
    for (int i = 0; i < n_devices; ++i)
        if (devices[i] < 0 || devices[i] >= count) return fail(WN_ERR_INVALID_ARGUMENT, "device %d out of range (%d devices)", devices[i], count);
    // One packed blob, copied device to device (NVLink peer copies where the topology allows, staged by the driver otherwise);
    // all copies are in flight together, then every replica validates and adopts its copy.
    std::vector<void*> staging((size_t)n_devices, nullptr);
    std::vector<cudaStream_t> streams((size_t)n_devices, nullptr);
    wn_status st = WN_OK;
    const size_t bytes = (size_t)src->hdr.total_bytes;
    for (int i = 0; i < n_devices && st == WN_OK; ++i) {
        DeviceGuard g(devices[i]);
        if (!g.ok) {
            st = fail(WN_ERR_CUDA, "cannot select CUDA device %d", devices[i]);
            break;
        }
        if (devices[i] != src->device) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], src->device) == cudaSuccess && can) {
                const cudaError_t pe = cudaDeviceEnablePeerAccess(src->device, 0);
                if (pe != cudaSuccess) cudaGetLastError(); // already enabled, or not permitted: the copy still works
            }
        }
        cudaError_t ce = cudaMalloc(&staging[i], bytes);
        if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking);
        if (ce == cudaSuccess) ce = cudaMemcpyPeerAsync(staging[i], devices[i], src->blob, src->device, bytes, streams[i]);
        if (ce != cudaSuccess) {
            cudaGetLastError();
            st = fail(ce == cudaErrorMemoryAllocation ? WN_ERR_OUT_OF_MEMORY : WN_ERR_CUDA, "replicating to device %d failed: %s", devices[i],
                      cudaGetErrorString(ce));
        }
    }
    for (int i = 0; i < n_devices; ++i) {
        if (!streams[i]) continue;
        DeviceGuard g(devices[i]);
        const cudaError_t ce = cudaStreamSynchronize(streams[i]);
        if (ce != cudaSuccess && st == WN_OK) st = fail(WN_ERR_CUDA, "replicating to device %d failed: %s", devices[i], cudaGetErrorString(ce));
        cudaStreamDestroy(streams[i]);
    }
    for (int i = 0; i < n_devices && st == WN_OK; ++i) {
        wn_options opt;
        wn_options_init(&opt);
        opt.device = devices[i];
        opt.accuracy_scale = 0.0f; // keep the source engine's
        DeviceGuard g(devices[i]);
        st = wn_create_from_packed(staging[i], (int64_t)bytes, &opt, &out[i]);
        if (st == WN_OK) out[i]->opt.accuracy_scale = out[i]->info.accuracy_scale = src->opt.accuracy_scale;
    }
    for (int i = 0; i < n_devices; ++i) {
        if (staging[i]) {
            DeviceGuard g(devices[i]);
            cudaFree(staging[i]);
        }
        if (st != WN_OK && out[i]) {
            wn_destroy(out[i]);
            out[i] = nullptr;
        }
    }
    return st;
}

wn_status wn_query_grid_multi(wn_engine* const* engines, int32_t n_engines, const float* origin, const float* spacing, const int64_t* dims,
                              float beta, uint32_t flags, float* out_omega, uint8_t* out_inside)
{
    if (!engines || n_engines < 1) return fail(WN_ERR_INVALID_ARGUMENT, "no engines");
    for (int i = 0; i < n_engines; ++i)
        if (!engines[i]) return fail(WN_ERR_INVALID_ARGUMENT, "engine %d is null", i);
    if (!out_omega && !out_inside) return fail(WN_ERR_INVALID_ARGUMENT, "no output requested");
    if (is_device_pointer(out_omega) || is_device_pointer(out_inside))
        return fail(WN_ERR_INVALID_ARGUMENT, "wn_query_grid_multi gathers into HOST buffers (each device holds only its own layers)");
    wn::GridDesc g;
    int64_t n = 0;
    wn_status s = check_grid(origin, spacing, dims, 0, dims ? dims[2] : 0, g, n);
    if (s != WN_OK) return s;
    if (n == 0) return WN_OK;
    if (n_engines == 1) return wn_query_grid(engines[0], origin, spacing, dims, 0, dims[2], beta, flags, out_omega, out_inside, nullptr);
    // Engine i takes rank i's share of the diagonal sharding (one y-part of every c-th tile layer: the same mix of work on every
    // device, no exchange step), one host thread per device; the compact per-device results are then laid out in lattice order.
    const bool bits = (flags & WN_QUERY_OUT_BITS) != 0 && out_inside;
    const int64_t nx = dims[0], ny = dims[1], nz = dims[2];
    std::vector<std::vector<float>> om((size_t)n_engines);
    std::vector<std::vector<uint8_t>> in((size_t)n_engines);
    std::vector<wn_status> status((size_t)n_engines, WN_OK);
    std::vector<std::string> errors((size_t)n_engines);
    std::vector<ShardLayout> lay((size_t)n_engines);
    std::vector<std::thread> pool;
    for (int i = 0; i < n_engines; ++i) {
        lay[i] = shard_layout(dims, i, n_engines);
        const int64_t ni = nx * lay[i].part_rows * lay[i].planes;
        if (out_omega) om[i].resize((size_t)ni);
        if (out_inside) in[i].resize((size_t)(bits ? (ni + 7) / 8 : ni));
        pool.emplace_back([&, i, ni]() {
            if (ni == 0) return;
            status[i] = wn_query_grid_sharded(engines[i], origin, spacing, dims, i, n_engines, beta, flags, out_omega ? om[i].data() : nullptr,
                                              out_inside ? in[i].data() : nullptr, nullptr);
            if (status[i] != WN_OK) errors[i] = wn_last_error();
        });
    }
    for (auto& t : pool) t.join();
    for (int i = 0; i < n_engines; ++i)
        if (status[i] != WN_OK) return fail(status[i], "device %d: %s", engines[i]->device, errors[i].c_str());
    for (int i = 0; i < n_engines; ++i) {
        const ShardLayout& L = lay[i];
        int64_t local = 0; // points of engine i consumed so far
        for (int64_t lz = i % L.c; lz * 8 < nz; lz += L.c) {
            int64_t q = ((i - lz) / L.c) % L.Q;
            if (q < 0) q += L.Q;
            const int64_t y0 = L.Q > 1 ? q * L.part_rows : 0, chunk = L.part_rows * nx; // one plane of the unit
            for (int64_t z = lz * 8; z < std::min<int64_t>(nz, lz * 8 + 8); ++z, local += chunk) {
                const int64_t first = (z * ny + y0) * nx;
                if (out_omega) memcpy(out_omega + first, om[i].data() + local, (size_t)chunk * sizeof(float));
                if (out_inside && bits) {
                    if ((first & 7) == 0 && (local & 7) == 0 && ((chunk & 7) == 0 || first + chunk == nx * ny * nz)) {
                        memcpy(out_inside + first / 8, in[i].data() + local / 8, (size_t)(chunk + 7) / 8);
                    } else { // ragged lattices only: bit by bit
                        for (int64_t k = 0; k < chunk; ++k) {
                            const int bit = (in[i][(size_t)((local + k) >> 3)] >> ((local + k) & 7)) & 1;
                            uint8_t& dst = out_inside[(first + k) >> 3];
                            dst = (uint8_t)((dst & ~(1u << ((first + k) & 7))) | (bit << ((first + k) & 7)));
                        }
                    }
                } else if (out_inside) {
                    memcpy(out_inside + first, in[i].data() + local, (size_t)chunk);
                }
            }
        }
    }
    return WN_OK;
}

wn_status wn_debug_last_plan(const wn_engine* e, int32_t* out, int64_t capacity_tiles, int64_t* num_tiles)
{
    if (!e || !num_tiles) return fail(WN_ERR_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(e->mu);
    *num_tiles = e->last_plan_tiles;
    if (!out) return WN_OK;
    if (capacity_tiles < e->last_plan_tiles) return fail(WN_ERR_INVALID_ARGUMENT, "output buffer too small");
    if (e->last_plan_tiles == 0) return WN_OK;
    DeviceGuard guard(e->device);
    std::vector<wn::TileHeader> h((size_t)e->last_plan_tiles);
    WN_CUDA(cudaDeviceSynchronize());
    WN_CUDA(cudaMemcpy(h.data(), (const char*)e->s_plan[0].hdr.p + 256, h.size() * sizeof(wn::TileHeader), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < h.size(); ++i) {
        out[4 * i + 0] = h[i].n_cond;
        out[4 * i + 1] = h[i].n_dir;
        out[4 * i + 2] = h[i].n_tri;
        out[4 * i + 3] = h[i].flags;
    }
    return WN_OK;
}

wn_status wn_debug_sort_pairs_u64(uint64_t* keys, uint32_t* values, int64_t n, int32_t begin_bit, int32_t end_bit)
{
    return debug_sort<uint64_t>(keys, values, n, begin_bit, end_bit);
}

wn_status wn_debug_sort_pairs_u32(uint32_t* keys, uint32_t* values, int64_t n, int32_t begin_bit, int32_t end_bit)
{
    return debug_sort<uint32_t>(keys, values, n, begin_bit, end_bit);
}

wn_status wn_debug_fma_peak(int32_t device, int32_t iters, float* tflops, float* ms_out)
{
    if (!tflops) return fail(WN_ERR_INVALID_ARGUMENT, "null argument");
    wn_options o;
    wn_options_init(&o);
    o.device = device;
    int dev = 0;
    wn_status s = resolve_device(&o, &dev);
    if (s != WN_OK) return s;
    DeviceGuard guard(dev);
    if (iters <= 0) iters = 4096;
    cudaDeviceProp prop;
    WN_CUDA(cudaGetDeviceProperties(&prop, dev));
    const int blocks = prop.multiProcessorCount * 8;
    float* sink = nullptr;
    WN_CUDA(cudaMalloc((void**)&sink, sizeof(float)));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    wn::k_fma_peak<<<blocks, 256>>>(iters / 8 + 1, sink); // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        wn::k_fma_peak<<<blocks, 256>>>(iters, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        best = std::min(best, ms);
    }
    cudaError_t ce = cudaGetLastError();
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(sink);
    WN_CUDA(ce);
    const double flops = (double)blocks * 256.0 * (double)iters * 64.0 * 2.0;
    *tflops = (float)(flops / (best * 1e-3) / 1e12);
    if (ms_out) *ms_out = best;
    return WN_OK;
}

} // extern "C"
