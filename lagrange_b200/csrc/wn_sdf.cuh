// wn_sdf.cuh — K10: narrow-band distance to the mesh on a cell-centred lattice, signed by the winding number.
// The GPU side of the reference's one production caller, volume::mesh_to_volume with Sign::WindingNumber
// (modules/volume/src/mesh_to_volume.cpp:147-183: OpenVDB meshToVolume, exterior = interior band width = 3 voxels, the
// interior test is FastWindingNumber::is_inside at every voxel centre). SURVEY.md section 8(f) N1/N3.
//
// Distance: closest-triangle search on the same packed hierarchy the winding-number queries walk. Every record already
// carries a bounding sphere of its subtree (expansion centre P, box-corner radius R: every triangle of the node lies within
// R of P), so a subtree is culled when |q - P| - R exceeds the best distance so far, which starts at the band width: only
// the hierarchy near the voxel is ever opened, and voxels outside the band stop after a few tests. Same warp-cooperative
// stackless walk as warp_traverse: the warp opens a subtree when any lane needs it.
#pragma once

#include "wn_query.cuh"

namespace wn {

struct SdfArgs
{
    WnTreeView tree;
    GridDesc g;
    int tiles_x, tiles_y;
    float band;            // world units; distances are clamped to it
    const uint8_t* inside; // [nz*ny*nx] winding-number sign per voxel, or null: unsigned distance
    float* out;            // [nz*ny*nx] signed distance (negative inside), +-band outside the narrow band
    unsigned long long* active; // optional: number of voxels with |d| < band
};

// CTA = 8x8x8 voxels, warp = 4x4x4, two voxels per lane (z and z+2), like k_query<2, GRID>.
__global__ void __launch_bounds__(kQueryThreads) k_sdf_grid(const SdfArgs a)
{
    constexpr int QPL = 2;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int bx = blockIdx.x % a.tiles_x, by = (blockIdx.x / a.tiles_x) % a.tiles_y, bz = blockIdx.x / (a.tiles_x * a.tiles_y);
    const int x = bx * 8 + (wid & 1) * 4 + (lane & 3);
    const int y = by * 8 + ((wid >> 1) & 1) * 4 + ((lane >> 2) & 3);
    float qx[QPL], qy[QPL], qz[QPL], best[QPL], best2[QPL];
    bool valid[QPL];
    int64_t oidx[QPL];
#pragma unroll
    for (int k = 0; k < QPL; ++k) {
        const int z = bz * 8 + (wid >> 2) * 4 + 2 * k + (lane >> 4);
        valid[k] = x < a.g.nx && y < a.g.ny && z < a.g.nz;
        qx[k] = wn_lattice_coord(a.g.ox, a.g.sx, x);
        qy[k] = wn_lattice_coord(a.g.oy, a.g.sy, y);
        qz[k] = wn_lattice_coord(a.g.oz, a.g.sz, z);
        oidx[k] = valid[k] ? ((int64_t)z * a.g.ny + y) * a.g.nx + x : -1;
        best[k] = a.band;
        best2[k] = a.band * a.band;
    }
    const float4* __restrict__ hot = a.tree.hot;
    const float4* __restrict__ tris = a.tree.tri;
    const int n = a.tree.n_entries;
    int i = 0;
    while (i < n) {
        const float4 f0 = __ldg(hot + 2 * (int64_t)i);
        const int lk = __float_as_int(__ldg(&hot[2 * (int64_t)i + 1].w));
        const bool leaf = __float_as_int(f0.w) < 0;
        const float R = sqrtf(fabsf(f0.w)); // +inf for records that are never approximated: never culled
        bool need[QPL], any = false;
#pragma unroll
        for (int k = 0; k < QPL; ++k) {
            const float rx = qx[k] - f0.x, ry = qy[k] - f0.y, rz = qz[k] - f0.z;
            const float d2 = rx * rx + ry * ry + rz * rz;
            const float reach = R + best[k];
            // conservative: a subtree is dropped only if it is clearly out of reach (1e-5 relative slack for rounding)
            need[k] = valid[k] && !(d2 > reach * reach * 1.00001f);
            any |= need[k];
        }
        if (!__any_sync(kFull, any)) {
            i = leaf ? i + 1 : lk;
            continue;
        }
        if (leaf) {
            const int first = lk >> WN_LEAF_COUNT_BITS, count = (lk & (WN_MAX_LEAF_SIZE - 1)) + 1;
            for (int tt = 0; tt < count; ++tt) {
                const float4 ta = __ldg(tris + 3 * (int64_t)(first + tt));
                const float4 tb = __ldg(tris + 3 * (int64_t)(first + tt) + 1);
                const float4 tc = __ldg(tris + 3 * (int64_t)(first + tt) + 2);
#pragma unroll
                for (int k = 0; k < QPL; ++k) {
                    if (need[k]) {
                        const float d2 = wn_point_tri_dist2(qx[k], qy[k], qz[k], ta, tb, tc);
                        if (d2 < best2[k]) {
                            best2[k] = d2;
                            best[k] = sqrtf(d2);
                        }
                    }
                }
            }
        }
        i = i + 1;
    }
    unsigned int act = 0;
#pragma unroll
    for (int k = 0; k < QPL; ++k) {
        if (oidx[k] >= 0) {
            const bool in = a.inside ? a.inside[oidx[k]] != 0 : false;
            a.out[oidx[k]] = in ? -best[k] : best[k];
            act += best[k] < a.band ? 1u : 0u;
        }
    }
    if (a.active) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) act += __shfl_xor_sync(kFull, act, o);
        if (lane == 0 && act) atomicAdd(a.active, (unsigned long long)act);
    }
}

// ---- K11: closest point on the mesh for arbitrary query points ---------------------------------------------------------------
// Replaces TriangleAABBTree::get_closest_point(p, triangle_id, closest_point, closest_sq_dist)
// (modules/bvh/include/lagrange/bvh/TriangleAABBTree.h:84-88; consumer modules/bvh/src/compute_mesh_distances.cpp:73), batched,
// on the winding-number engine's own packed hierarchy (every record's (P, R) is a bounding sphere of its subtree).
// Phase 1, per lane: greedy descent to the child whose sphere is nearest gives a first triangle and with it a finite search
// radius (the reference seeds its search with a hint triangle for the same reason). Phase 2, warp-cooperative: the stackless
// walk of k_sdf_grid, a subtree is opened when any lane can still improve.
struct ClosestArgs
{
    WnTreeView tree;
    const unsigned* tri_order; // depth-first position -> input triangle id
    const float* q;            // [n * 3]
    const unsigned* perm;      // optional Morton order of the queries
    int64_t n;
    float max_dist;            // > 0: search radius (points farther away report it, triangle -1); <= 0: unbounded
    float* out_sqdist;         // [n] or null
    int* out_tri;              // [n] or null
    float* out_xyz;            // [n * 3] or null
};

__global__ void __launch_bounds__(kQueryThreads) k_closest_point(const ClosestArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t slot = (int64_t)blockIdx.x * kQueryThreads + threadIdx.x;
    const bool valid = slot < a.n;
    const int64_t p = valid ? (a.perm ? (int64_t)a.perm[slot] : slot) : 0;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (valid) {
        qx = __ldg(a.q + 3 * p);
        qy = __ldg(a.q + 3 * p + 1);
        qz = __ldg(a.q + 3 * p + 2);
    }
    const float4* __restrict__ hot = a.tree.hot;
    const float4* __restrict__ tris = a.tree.tri;
    const int n = a.tree.n_entries;
    const bool bounded = a.max_dist > 0.0f;
    float best2 = bounded ? a.max_dist * a.max_dist : 3.0e38f, best = bounded ? a.max_dist : 1.8e19f;
    int best_pos = -1;
    float cx = qx, cy = qy, cz = qz;
    auto try_leaf = [&](int lk) {
        const int first = lk >> WN_LEAF_COUNT_BITS, count = (lk & (WN_MAX_LEAF_SIZE - 1)) + 1;
        for (int tt = 0; tt < count; ++tt) {
            float x, y, z;
            const float d2 = wn_point_tri_closest(qx, qy, qz, __ldg(tris + 3 * (int64_t)(first + tt)), __ldg(tris + 3 * (int64_t)(first + tt) + 1),
                                                  __ldg(tris + 3 * (int64_t)(first + tt) + 2), x, y, z);
            if (d2 < best2) {
                best2 = d2;
                best = sqrtf(d2);
                best_pos = first + tt;
                cx = x, cy = y, cz = z;
            }
        }
    };
    // ---- phase 1: greedy descent (divergent, ~depth dependent loads per lane) ----------------------------------------------
    if (valid && n > 0 && !bounded) {
        int i = 0;
        for (int guard = 0; guard < 128; ++guard) {
            const float4 f0 = __ldg(hot + 2 * (int64_t)i);
            if (__float_as_int(f0.w) < 0) {
                try_leaf(__float_as_int(__ldg(&hot[2 * (int64_t)i + 1].w)));
                break;
            }
            const int4 k4 = __ldg(a.tree.kids + i);
            const int kid[4] = {k4.x, k4.y, k4.z, k4.w};
            int next = -1;
            float bestgap = 3.4e38f;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (kid[s] < 0) continue;
                const float4 g0 = __ldg(hot + 2 * (int64_t)kid[s]);
                const float rx = qx - g0.x, ry = qy - g0.y, rz = qz - g0.z;
                const float R = fabsf(g0.w);
                // distance to the child's bounding sphere; records that are never approximated carry R = inf: rank them by centre distance
                const float d = sqrtf(rx * rx + ry * ry + rz * rz);
                const float gap = R < 3.0e38f ? d - sqrtf(R) : d;
                if (gap < bestgap) {
                    bestgap = gap;
                    next = kid[s];
                }
            }
            if (next < 0) break;
            i = next;
        }
    }
    // ---- phase 2: warp-cooperative culling walk -----------------------------------------------------------------------------
    int i = 0;
    while (i < n) {
        const float4 f0 = __ldg(hot + 2 * (int64_t)i);
        const int lk = __float_as_int(__ldg(&hot[2 * (int64_t)i + 1].w));
        const bool leaf = __float_as_int(f0.w) < 0;
        const float R = sqrtf(fabsf(f0.w));
        const float rx = qx - f0.x, ry = qy - f0.y, rz = qz - f0.z;
        const float d2 = rx * rx + ry * ry + rz * rz;
        const float reach = R + best;
        const bool need = valid && !(d2 > reach * reach * 1.00001f);
        if (!__any_sync(kFull, need)) {
            i = leaf ? i + 1 : lk;
            continue;
        }
        if (leaf && need) try_leaf(lk);
        i = i + 1;
    }
    (void)lane;
    if (valid) {
        if (a.out_sqdist) a.out_sqdist[p] = best2;
        if (a.out_tri) a.out_tri[p] = best_pos >= 0 ? (int)__ldg(a.tri_order + best_pos) : -1;
        if (a.out_xyz) {
            a.out_xyz[3 * p] = cx;
            a.out_xyz[3 * p + 1] = cy;
            a.out_xyz[3 * p + 2] = cz;
        }
    }
}

// ---- sparse narrow band: compaction of the active voxels of a dense signed-distance block -----------------------------------------
// (what an OpenVDB grid keeps: modules/volume/src/mesh_to_volume.cpp:160-183 returns a FloatGrid whose active voxels are the band)
constexpr int kCompactItems = 8; // voxels per thread
__global__ void __launch_bounds__(kQueryThreads) k_band_count(const float* __restrict__ sdf, int64_t n, float band, uint32_t* __restrict__ counts)
{
    const int64_t base = (int64_t)blockIdx.x * (kQueryThreads * kCompactItems);
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < kCompactItems; ++k) {
        const int64_t i = base + (int64_t)k * kQueryThreads + threadIdx.x;
        c += (i < n && fabsf(sdf[i]) < band) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
    __shared__ unsigned s[kQueryWarps];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < kQueryWarps; ++w) t += s[w];
        counts[blockIdx.x] = t;
    }
}
// sign bits of a dense signed-distance block: voxel i -> bit (i & 7) of out[i >> 3] (1 = inside = negative distance)
__global__ void __launch_bounds__(256) k_sign_bits(const float* __restrict__ sdf, int64_t n, uint8_t* __restrict__ out)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b * 8 >= n) return;
    unsigned v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int64_t i = b * 8 + k;
        if (i < n && sdf[i] < 0.0f) v |= 1u << k;
    }
    out[b] = (uint8_t)v;
}
// offsets = exclusive scan of counts. Output is ordered by linear voxel index.
__global__ void __launch_bounds__(kQueryThreads) k_band_scatter(const float* __restrict__ sdf, int64_t n, float band, const uint32_t* __restrict__ offsets,
                                                                int64_t capacity, int64_t* __restrict__ out_index, float* __restrict__ out_value)
{
    __shared__ unsigned s_warp[kQueryWarps];
    __shared__ unsigned s_run;
    const int64_t base = (int64_t)blockIdx.x * (kQueryThreads * kCompactItems);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = offsets[blockIdx.x];
    __syncthreads();
    for (int k = 0; k < kCompactItems; ++k) {
        const int64_t i = base + (int64_t)k * kQueryThreads + threadIdx.x;
        const float v = i < n ? sdf[i] : 0.0f;
        const bool act = i < n && fabsf(v) < band;
        const unsigned m = __ballot_sync(kFull, act);
        if (lane == 0) s_warp[wid] = __popc(m);
        __syncthreads();
        unsigned before = s_run;
        for (int w = 0; w < wid; ++w) before += s_warp[w];
        const int64_t dst = (int64_t)before + __popc(m & ((1u << lane) - 1u));
        if (act && dst < capacity) {
            out_index[dst] = i;
            out_value[dst] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned t = 0;
            for (int w = 0; w < kQueryWarps; ++w) t += s_warp[w];
            s_run += t;
        }
        __syncthreads();
    }
}

} // namespace wn
