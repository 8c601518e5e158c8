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

// ---- K3'': SAH-guided split positions (WN_HIERARCHY_KD_SAH). Same idea as K3', but a range sorted along its axis is cut
//      where  area(left) n_left + area(right) n_right  is least among the boundaries of 16 bins of equal width along its axis
//      (exact triangle boxes and centroid counts per bin, the reference builder's rule), so node ranges are explicit: per level, arrays over the nodes in position order.
struct KdxLevel
{
    const unsigned* start; // [count + 1], start[count] = N
    const int* pid;        // [count] id of the parent internal node (-1: root)
    const unsigned char* meta; // [count] bit 0: slot in the parent, bit 1: already linked into the tree arrays
    int count;
};

__global__ void __launch_bounds__(kBuildThreads) k_kdx_bounds(const float* __restrict__ v, const int* __restrict__ tri,
                                                              const unsigned* __restrict__ perm, const unsigned* __restrict__ node_of, int N,
                                                              const unsigned* __restrict__ start, int leaf, int* __restrict__ bounds)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned node = 0xffffffffu;
    int enc[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
    bool active = false;
    if (p < N) {
        const unsigned i = node_of[p];
        active = (int)(start[i + 1] - start[i]) > leaf;
        if (active) {
            node = i;
            float c[3];
            kd_centroid(v, tri, perm[p], c);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (c[a] == c[a]) {
                    enc[a] = float_to_ordered(c[a]);
                    enc[3 + a] = enc[a];
                }
            }
        }
    }
    const unsigned same = __match_any_sync(0xffffffffu, node);
    if (same == 0xffffffffu) {
        if (!active) return;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            enc[a] = __reduce_min_sync(0xffffffffu, enc[a]);
            enc[3 + a] = __reduce_max_sync(0xffffffffu, enc[3 + a]);
        }
        if ((threadIdx.x & 31) != 0) return;
    } else if (!active) {
        return;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        atomicMin(&bounds[6 * (size_t)node + a], enc[a]);
        atomicMax(&bounds[6 * (size_t)node + 3 + a], enc[3 + a]);
    }
}

__global__ void __launch_bounds__(kBuildThreads) k_kdx_keys(const float* __restrict__ v, const int* __restrict__ tri,
                                                            const unsigned* __restrict__ perm, const unsigned* __restrict__ node_of, int N,
                                                            const unsigned* __restrict__ start, int leaf, const int* __restrict__ bounds,
                                                            uint64_t* __restrict__ keys)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const unsigned i = node_of[p];
    unsigned q = 0;
    if ((int)(start[i + 1] - start[i]) > leaf) {
        const int* b = bounds + 6 * (size_t)i;
        float blo[3], ext[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            blo[a] = ordered_to_float(b[a]);
            ext[a] = b[3 + a] >= b[a] ? WN_SUB(ordered_to_float(b[3 + a]), blo[a]) : 0.0f;
        }
        const int axis = wn_kd_axis(ext);
        float c[3];
        kd_centroid(v, tri, perm[p], c);
        q = wn_kd_quant(c[axis], blo[axis], ext[axis]);
    }
    keys[p] = ((uint64_t)i << 16) | q;
}

// Bin table of every range with at least WN_KDX_MIN_SAH triangles: triangle boxes and centroid counts of WN_KDX_BINS bins of
// equal width along the range's split axis; row = start / WN_KDX_MIN_SAH, WN_KDX_ROW ints per row. Consecutive positions are
// sorted along that axis, so a warp mostly hits one (row, bin): it reduces first and issues one set of atomics (near the root
// every warp does; without this the root's few addresses would take 1.3 M atomics each).
__global__ void __launch_bounds__(kBuildThreads) k_kdx_segboxes(const float* __restrict__ v, const int* __restrict__ tri,
                                                                const unsigned* __restrict__ perm, const unsigned* __restrict__ node_of, int N,
                                                                const unsigned* __restrict__ start, int leaf, const int* __restrict__ bounds,
                                                                int* __restrict__ segbox)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    long long slot = -1; // row * BINS + bin
    int enc[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
    if (p < N) {
        const unsigned i = node_of[p];
        const int s = (int)start[i], n = (int)(start[i + 1] - start[i]);
        if (n > leaf && n >= WN_KDX_MIN_SAH) {
            const int* b = bounds + 6 * (size_t)i;
            float blo[3], ext[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                blo[a] = ordered_to_float(b[a]);
                ext[a] = b[3 + a] >= b[a] ? WN_SUB(ordered_to_float(b[3 + a]), blo[a]) : 0.0f;
            }
            const int axis = wn_kd_axis(ext);
            const unsigned t = perm[p];
            float c[3];
            kd_centroid(v, tri, t, c);
            const int bin = wn_kdx_bin(wn_kd_quant(c[axis], blo[axis], ext[axis]));
            slot = (long long)(s / WN_KDX_MIN_SAH) * WN_KDX_BINS + bin;
            const int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float x0 = v[3 * i0 + a], x1 = v[3 * i1 + a], x2 = v[3 * i2 + a];
                const float lo = fminf(x0, fminf(x1, x2)), hi = fmaxf(x0, fmaxf(x1, x2)); // NaN coordinates are ignored by fminf/fmaxf
                if (lo == lo) enc[a] = float_to_ordered(lo);
                if (hi == hi) enc[3 + a] = float_to_ordered(hi);
            }
        }
    }
    int add = 1;
    const unsigned same = __match_any_sync(0xffffffffu, slot);
    if (same == 0xffffffffu) {
        if (slot < 0) return;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            enc[a] = __reduce_min_sync(0xffffffffu, enc[a]);
            enc[3 + a] = __reduce_max_sync(0xffffffffu, enc[3 + a]);
        }
        if ((threadIdx.x & 31) != 0) return;
        add = 32;
    } else if (slot < 0) {
        return;
    }
    const long long row = slot / WN_KDX_BINS;
    const int bin = (int)(slot % WN_KDX_BINS);
    int* r = segbox + (size_t)row * WN_KDX_ROW;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (enc[a] != INT_MAX) atomicMin(&r[bin * 6 + a], enc[a]);
        if (enc[3 + a] != INT_MIN) atomicMax(&r[bin * 6 + 3 + a], enc[3 + a]);
    }
    atomicAdd(&r[WN_KDX_BINS * 6 + bin], add);
}

// empty bin table for the ranges that will be measured at this level (one thread per range and value)
__global__ void __launch_bounds__(kBuildThreads) k_kdx_init_segbox_nodes(const unsigned* __restrict__ start, int count, int leaf,
                                                                         int* __restrict__ segbox)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)(g / WN_KDX_ROW), k = (int)(g % WN_KDX_ROW);
    if (i >= count) return;
    const int s = (int)start[i], n = (int)(start[i + 1] - start[i]);
    if (n <= leaf || n < WN_KDX_MIN_SAH) return;
    segbox[(size_t)(s / WN_KDX_MIN_SAH) * WN_KDX_ROW + k] = k >= WN_KDX_BINS * 6 ? 0 : ((k % 6) < 3 ? INT_MAX : INT_MIN);
}

// one thread per node: left count (0 = not split at this level), children count for the scan, bookkeeping in res[]:
// res[0] += nodes split, res[2] = root gap (level 0)
__global__ void __launch_bounds__(kBuildThreads) k_kdx_split(KdxLevel L, int N, int leaf, int level, const int* __restrict__ segbox,
                                                             int* __restrict__ nl, uint32_t* __restrict__ cnt, int* __restrict__ res)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.count) return;
    const int s = (int)L.start[i], n = (int)(L.start[i + 1] - L.start[i]);
    int left = 0;
    if (n > leaf) {
        left = n / 2;
        if (n >= WN_KDX_MIN_SAH) {
            float box[WN_KDX_BINS * 6];
            int bcnt[WN_KDX_BINS];
            const int* row = segbox + (size_t)(s / WN_KDX_MIN_SAH) * WN_KDX_ROW;
            for (int k = 0; k < WN_KDX_BINS * 6; ++k) // untouched sentinels decode to an empty box, like the host emulation's
                box[k] = row[k] == INT_MAX ? 3.4e38f : (row[k] == INT_MIN ? -3.4e38f : ordered_to_float(row[k]));
            for (int k = 0; k < WN_KDX_BINS; ++k) bcnt[k] = row[WN_KDX_BINS * 6 + k];
            left = wn_kdx_choose(n, box, bcnt);
        }
        atomicAdd(&res[0], 1);
        atomicMax(&res[3], max(left, n - left)); // largest range of the next level: the host skips the SAH passes below the threshold
    }
    nl[i] = left;
    cnt[i] = left > 0 ? 2u : 1u;
    if (level == 0) res[2] = left > 0 ? s + left - 1 : N / 2 - 1;
}

// one thread per node: link it into the tree arrays, write its children (or itself, carried) into the next level
__global__ void __launch_bounds__(kBuildThreads) k_kdx_scatter(KdxLevel L, int N, int level, const int* __restrict__ nl, const uint32_t* __restrict__ off,
                                                               const int* __restrict__ res, unsigned* __restrict__ nstart, int* __restrict__ npid,
                                                               unsigned char* __restrict__ nmeta, int* __restrict__ child, int* __restrict__ parent,
                                                               unsigned char* __restrict__ slot, unsigned char* __restrict__ skip, int* __restrict__ total)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.count) return;
    const int s = (int)L.start[i], n = (int)(L.start[i + 1] - L.start[i]);
    const int pid = L.pid[i], side = L.meta[i] & 1;
    const bool linked = (L.meta[i] & 2) != 0;
    const int root_gap = res[2], nI = N - 1;
    const int left = nl[i];
    const unsigned j = off[i];
    if (left > 0) {
        const int id = wn_kdx_gap_id(s + left - 1, root_gap);
        if (pid >= 0) {
            child[2 * (size_t)pid + side] = id;
            parent[id] = pid;
            slot[id] = (unsigned char)side;
        }
        if (skip) skip[id] = (level & 1) ? 1 : 0;
        const int cs[2] = {s, s + left}, cm[2] = {left, n - left};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            nstart[j + c] = (unsigned)cs[c];
            npid[j + c] = id;
            if (cm[c] == 1) { // a single triangle: leaf node nI + position, linked right away
                child[2 * (size_t)id + c] = nI + cs[c];
                parent[nI + cs[c]] = id;
                slot[nI + cs[c]] = (unsigned char)c;
                nmeta[j + c] = (unsigned char)(c | 2);
            } else {
                nmeta[j + c] = (unsigned char)c;
            }
        }
    } else {
        if (!linked) wn_kdx_emit_halving(N, s, n, pid, side, root_gap, child, parent, slot, skip);
        nstart[j] = (unsigned)s;
        npid[j] = pid;
        nmeta[j] = (unsigned char)(side | 2);
    }
    if (i == L.count - 1) {
        const int tot = (int)j + (left > 0 ? 2 : 1);
        nstart[tot] = (unsigned)N;
        *total = tot;
    }
}

// elements follow their node into the next level
__global__ void __launch_bounds__(kBuildThreads) k_kdx_assign(unsigned* __restrict__ node_of, int N, const unsigned* __restrict__ start,
                                                              const int* __restrict__ nl, const uint32_t* __restrict__ off)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const unsigned i = node_of[p];
    const int left = nl[i];
    node_of[p] = off[i] + ((left > 0 && p >= (int)start[i] + left) ? 1u : 0u);
}

__global__ void __launch_bounds__(kBuildThreads) k_fill_int(int* __restrict__ p, int64_t n, int value)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = value;
}

} // namespace wn
