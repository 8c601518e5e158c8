// wn_build.cuh — sm_100a kernels of the hierarchy build and their host-side orchestration.
//
// Replaces UT_SolidAngle::init (called at modules/winding/src/FastWindingNumber.cpp:57): triangle boxes + UT_BVH<4>
// SAH build + post-order moment pass on the CPU become
//   K1 k_centroid_bounds / k_morton   scene bounds of the triangle centroids, 63- or 30-bit Morton codes
//   K2 wn::radix_sort_pairs           (wn_sort.cuh)
//   K3 k_lbvh                         Karras 2012 binary radix tree, one thread per internal node
//      k_import_topology              (oracle-tree mode) caller-supplied topology instead of K1-K3
//   K4 k_moments_climb                per-triangle moments + bottom-up merge with arrival counters (A.2, A.3)
//      k_vertex_radius                optional exact bounding radius about each node's centroid
//   K5 k_pack                         depth-first index + skip link by walking up, folded 6 x float4 records
// All of these are HBM/latency bound integer + a little FP32 work: one thread per element, coalesced SoA arrays,
// grids sized to the element count (>> 148 SMs x resident CTAs for the configs' sizes).
#pragma once

#include <cuda_runtime.h>

#include "wn_build_core.cuh"
#include "wn_sort.cuh"

namespace wn {

constexpr int kBuildThreads = 256;

__device__ __forceinline__ int float_to_ordered(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float ordered_to_float(int i)
{
    const int j = i >= 0 ? i : i ^ 0x7fffffff;
#if defined(__CUDA_ARCH__)
    return __int_as_float(j);
#else
    float f;
    memcpy(&f, &j, 4);
    return f;
#endif
}

// bounds[0..2] = min (ordered ints), bounds[3..5] = max. Initialised to INT_MAX / INT_MIN by the host.
__global__ void __launch_bounds__(kBuildThreads) k_centroid_bounds(const float* __restrict__ v, const int* __restrict__ tri, int nT,
                                                                   int nV, int* __restrict__ bounds, int* __restrict__ err)
{
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nT; t += gridDim.x * blockDim.x) {
        const int i0 = tri[3 * (int64_t)t], i1 = tri[3 * (int64_t)t + 1], i2 = tri[3 * (int64_t)t + 2];
        if ((unsigned)i0 >= (unsigned)nV || (unsigned)i1 >= (unsigned)nV || (unsigned)i2 >= (unsigned)nV) {
            *err = 3; // vertex index out of range
            continue;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float c = wn_centroid_coord(v[3 * (int64_t)i0 + a], v[3 * (int64_t)i1 + a], v[3 * (int64_t)i2 + a]);
            if (c == c) { // ignore NaN centroids
                lo[a] = fminf(lo[a], c);
                hi[a] = fmaxf(hi[a], c);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&bounds[a], float_to_ordered(lo[a]));
            atomicMax(&bounds[3 + a], float_to_ordered(hi[a]));
        }
    }
}

template <typename K>
__global__ void __launch_bounds__(kBuildThreads) k_morton(const float* __restrict__ v, const int* __restrict__ tri, int nT,
                                                          const int* __restrict__ bounds, int bits_per_axis, K* __restrict__ keys,
                                                          unsigned* __restrict__ vals)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nT) return;
    const float lx = ordered_to_float(bounds[0]), ly = ordered_to_float(bounds[1]), lz = ordered_to_float(bounds[2]);
    const float ex = WN_SUB(ordered_to_float(bounds[3]), lx), ey = WN_SUB(ordered_to_float(bounds[4]), ly), ez = WN_SUB(ordered_to_float(bounds[5]), lz);
    const float ext = fmaxf(ex, fmaxf(ey, ez)); // cubic cells: one scale for all axes
    const float inv = ext > 0.0f ? WN_DIV(1.0f, ext) : 0.0f;
    const int i0 = tri[3 * (int64_t)t], i1 = tri[3 * (int64_t)t + 1], i2 = tri[3 * (int64_t)t + 2];
    const float cx = wn_centroid_coord(v[3 * (int64_t)i0], v[3 * (int64_t)i1], v[3 * (int64_t)i2]);
    const float cy = wn_centroid_coord(v[3 * (int64_t)i0 + 1], v[3 * (int64_t)i1 + 1], v[3 * (int64_t)i2 + 1]);
    const float cz = wn_centroid_coord(v[3 * (int64_t)i0 + 2], v[3 * (int64_t)i1 + 2], v[3 * (int64_t)i2 + 2]);
    keys[t] = (K)wn_morton(wn_unit_coord(cx, lx, inv), wn_unit_coord(cy, ly, inv), wn_unit_coord(cz, lz, inv), bits_per_axis);
    vals[t] = (unsigned)t;
}

__global__ void __launch_bounds__(kBuildThreads) k_widen_keys(const uint32_t* __restrict__ in, uint64_t* __restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

__global__ void __launch_bounds__(kBuildThreads) k_lbvh(const uint64_t* __restrict__ keys, int n, int* __restrict__ child,
                                                        int* __restrict__ parent, unsigned char* __restrict__ slot)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) wn_lbvh_node(keys, n, i, child, parent, slot);
}

__global__ void __launch_bounds__(kBuildThreads) k_import_topology(WnBuild b, const int* __restrict__ child_in, int* __restrict__ seen)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b.nI) wn_import_node(b, child_in, seen, i);
}
__global__ void __launch_bounds__(kBuildThreads) k_import_check(WnBuild b, const int* __restrict__ seen)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b.nI + b.nL) wn_import_check(b, seen, i);
}
__global__ void __launch_bounds__(kBuildThreads) k_iota(unsigned* __restrict__ p, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (unsigned)i;
}

__global__ void __launch_bounds__(kBuildThreads) k_moments_climb(WnBuild b)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < b.nL) wn_climb_leaf(b, l);
}
__global__ void __launch_bounds__(kBuildThreads) k_vertex_radius(WnBuild b)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < b.nL) wn_vertex_radius_leaf(b, l);
}
__global__ void __launch_bounds__(kBuildThreads) k_pack(WnBuild b)
{
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node < b.nI + b.nL) wn_pack_node(b, node);
}

__global__ void __launch_bounds__(kBuildThreads) k_ref23(WnBuild b, int64_t first, int64_t count, float* __restrict__ out)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    WnLocal d;
    wn_load_local(b.local + (first + k) * 9, d, false);
    float o[23];
    wn_local_to_ref23(d, o);
    for (int j = 0; j < 23; ++j) out[k * 23 + j] = o[j];
}

// Structural check of an adopted packed tree (wn_create_from_packed): every index the traversal, the tile planner or the
// distance query will dereference must stay inside the arrays.
__global__ void __launch_bounds__(kBuildThreads) k_validate_packed(WnTreeView t, int* __restrict__ err)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.n_entries) return;
    const float4 f0 = t.hot[2 * (int64_t)i], f1 = t.hot[2 * (int64_t)i + 1];
    const int lk = __float_as_int(f1.w);
    const bool leaf = __float_as_int(f0.w) < 0;
    bool bad = false;
    if (leaf) {
        const int first = lk >> WN_LEAF_COUNT_BITS, count = (lk & (WN_MAX_LEAF_SIZE - 1)) + 1;
        bad = lk < 0 || first < 0 || first + count > t.n_tris;
    } else {
        bad = lk <= i || lk > t.n_entries; // skip link: first entry after the subtree
        const int4 k = t.kids[i];
        const int kid[4] = {k.x, k.y, k.z, k.w};
        for (int s = 0; s < 4; ++s) bad = bad || kid[s] < -1 || (kid[s] >= 0 && (kid[s] <= i || kid[s] >= lk || kid[s] >= t.n_entries));
        bad = bad || (t.n_entries > 1 && i + 1 >= t.n_entries); // an internal entry has at least one entry below it
    }
    if (bad) *err = 1;
}

// WN_QUERY_OUT_BITS: in[i] in {0, 1}, i < n  ->  out[i >> 3] bit (i & 7). One thread per output byte; `in` is 8-byte aligned.
__global__ void __launch_bounds__(256) k_pack_bits(const uint8_t* __restrict__ in, int64_t n, uint8_t* __restrict__ out)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b * 8 >= n) return;
    unsigned long long x;
    if (b * 8 + 8 <= n) {
        x = *reinterpret_cast<const unsigned long long*>(in + b * 8);
    } else {
        x = 0;
        for (int k = 0; b * 8 + k < n; ++k) x |= (unsigned long long)in[b * 8 + k] << (8 * k);
    }
    out[b] = (uint8_t)((x * 0x0102040810204080ull) >> 56); // bytes are 0/1: byte k lands on bit k
}

inline int grid_for(int64_t n, int threads = kBuildThreads)
{
    return (int)((n + threads - 1) / threads);
}

} // namespace wn
