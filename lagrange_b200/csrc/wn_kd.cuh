// wn_kd.cuh — K3': balanced k-d hierarchy (object-median splits along the longest axis of the centroid bounds).
//
// Why: the records' acceptance radius is the box-corner radius about the area-weighted centroid, so the traversal does the
// least work on compact patches of equal triangle count. Morton cells (K3, Karras) cut a surface into slivers of very
// different sizes; on BASELINE cfg2 the query kernels run 13 % faster on this hierarchy than on the LBVH, *measured*
// (tools/median_experiment.py; the reference's own top-down SAH tree, imported, is +24 %), for ~4 ms more build time.
//
// How: the tree over the final triangle order is the implicit balanced one — a range of n triangles splits into its first
// n/2 and the rest — so a position's node at any level is a function of (N, position) alone and nothing but the order has
// to be computed: for level = 0, 1, ...: per-node centroid bounds (atomics, warp-aggregated), key = (node path << 16) |
// 16-bit coordinate along the node's longest axis, one stable radix sort (K2) of (key, triangle). log2(N) rounds.
// k_kd_tree then writes the same child / parent / slot arrays k_lbvh writes, and the rest of the build (K4, K5) is shared.
#pragma once

#include "wn_build.cuh"

namespace wn {

__global__ void __launch_bounds__(kBuildThreads) k_kd_init_bounds(int* __restrict__ bounds, int nodes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nodes * 6) bounds[i] = (i % 6) < 3 ? INT_MAX : INT_MIN;
}

__device__ __forceinline__ void kd_centroid(const float* __restrict__ v, const int* __restrict__ tri, unsigned t, float c[3])
{
    const int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
#pragma unroll
    for (int a = 0; a < 3; ++a) c[a] = wn_centroid_coord(v[3 * i0 + a], v[3 * i1 + a], v[3 * i2 + a]);
}

// bounds[node][0..2] = min, [3..5] = max of the centroids of the node's triangles (ordered ints)
__global__ void __launch_bounds__(kBuildThreads) k_kd_bounds(const float* __restrict__ v, const int* __restrict__ tri,
                                                             const unsigned* __restrict__ perm, int N, int level, int* __restrict__ bounds)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = p < N;
    int lo = 0, n = 0;
    unsigned path = 0xffffffffu;
    int enc[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
    bool active = false;
    if (in) {
        wn_kd_locate(N, p, level, lo, n, path);
        active = n >= 2;
        if (active) {
            float c[3];
            kd_centroid(v, tri, perm[p], c);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (c[a] == c[a]) { // NaN centroids do not take part
                    enc[a] = float_to_ordered(c[a]);
                    enc[3 + a] = enc[a];
                }
            }
        }
    }
    const unsigned key = active ? path : 0xffffffffu;
    const unsigned same = __match_any_sync(0xffffffffu, key);
    if (same == 0xffffffffu) {
        // the whole warp sits in one node (always true near the root): one set of atomics per warp
        if (!active) return;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            enc[a] = __reduce_min_sync(0xffffffffu, enc[a]);
            enc[3 + a] = __reduce_max_sync(0xffffffffu, enc[3 + a]);
        }
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                atomicMin(&bounds[6 * (size_t)path + a], enc[a]);
                atomicMax(&bounds[6 * (size_t)path + 3 + a], enc[3 + a]);
            }
        }
    } else if (active) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&bounds[6 * (size_t)path + a], enc[a]);
            atomicMax(&bounds[6 * (size_t)path + 3 + a], enc[3 + a]);
        }
    }
}

// key = (path << 16) | 16-bit position of the centroid along the longest axis of its node's centroid bounds
__global__ void __launch_bounds__(kBuildThreads) k_kd_keys(const float* __restrict__ v, const int* __restrict__ tri,
                                                           const unsigned* __restrict__ perm, int N, int level, const int* __restrict__ bounds,
                                                           uint64_t* __restrict__ keys, unsigned* __restrict__ vals)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    int lo, n;
    unsigned path;
    wn_kd_locate(N, p, level, lo, n, path);
    const unsigned t = perm[p];
    unsigned q = 0;
    if (n >= 2) {
        const int* b = bounds + 6 * (size_t)path;
        float blo[3], ext[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            blo[a] = ordered_to_float(b[a]);
            ext[a] = b[3 + a] >= b[a] ? WN_SUB(ordered_to_float(b[3 + a]), blo[a]) : 0.0f;
        }
        const int axis = wn_kd_axis(ext);
        float c[3];
        kd_centroid(v, tri, t, c);
        q = wn_kd_quant(c[axis], blo[axis], ext[axis]);
    }
    keys[p] = ((uint64_t)path << 16) | q;
    if (vals) vals[p] = t; // null: the values already are the permutation (sorted in place by the caller's buffers)
}

// One thread per internal node (gap): find its range by descending from the root, emit children / parents / slots in the
// layout of k_lbvh (internal nodes 0..N-2, leaf of sorted position p = node N-1 + p).
__global__ void __launch_bounds__(kBuildThreads) k_kd_tree(int N, int* __restrict__ child, int* __restrict__ parent,
                                                           unsigned char* __restrict__ slot, unsigned char* __restrict__ skip, int leaf_size)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N - 1) return;
    wn_kd_emit_node(N, g, child, parent, slot, skip, leaf_size);
}

} // namespace wn
