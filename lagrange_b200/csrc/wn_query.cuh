// wn_query.cuh — sm_100a tree query kernels (K6). Replace UT_SolidAngle::computeSolidAngle
// (modules/winding/src/FastWindingNumber.cpp:66,75) for whole batches.
//
// warp_traverse  : the traversal every kernel shares. One warp owns 32*QPL spatially adjacent queries and walks a
//                  depth-first sequence of records once for all of them (stackless: "descend" = next record, "skip
//                  subtree" = the record's skip link). Each lane keeps the reference's per-point semantics through a
//                  private resume index: a lane that accepted a far-field record ignores records until the end of that
//                  subtree, lanes that must descend keep going; the warp leaves a subtree only when no lane needs it.
//                  The accept test |q-P|^2 <= beta^2 R^2 is formed unfused, exactly like the reference, so every point
//                  takes the same branches as in the CPU algorithm. Far field = folded order-2 Taylor record, leaves =
//                  exact Van Oosterom-Strackee triangles. Record reads are warp-uniform float4 broadcasts.
// k_query        : generic kernel: the record sequence is the whole packed tree (small or incoherent batches, fallback).
// k_tile_plan +  : tiled path for coherent batches (lattices, Morton-sorted point sets). A tile is 8x8x8 lattice points
// k_tile_query     (or 512 consecutive sorted points). k_tile_plan classifies tree nodes against the tile's bounding
//                  sphere, breadth first, with all 256 threads:
//                    far for every point and no ancestor that some point accepted  -> "far set": its field is smooth over
//                        the tile, so it is evaluated once per tile at 4^3 Chebyshev points instead of once per query
//                    near for every point, internal                               -> dropped (nobody tests it), children expanded
//                    anything else                                                 -> item of the tile's record list
//                  then sorts the list back into depth-first order (bitonic, shared memory) and rebuilds skip links in
//                  list coordinates. k_tile_query runs warp_traverse over that short list (staged in shared memory) and
//                  adds the tensor-product Chebyshev interpolant of the far set. Which records a point accepts is
//                  unchanged; only where their sum is evaluated differs (interpolation error ~1e-5 * 4 pi, measured in
//                  tests). Executed work drops ~4x on the 512^3 / 1.3M-triangle configuration.
// All FP32 CUDA-core work (FMA pipe + MUFU rsqrt/atan): no tensor cores by design (BASELINE.json north_star).
#pragma once

#include <cuda_runtime.h>

#include "wn_exact.cuh"

namespace wn {

constexpr int kQueryThreads = 256;
constexpr int kQueryWarps = kQueryThreads / 32;
constexpr int kTileQPL = 2;                       // queries per lane in the tiled kernels: 8 warps * 32 * 2 = 512 = 8^3
constexpr int kTileQueries = kQueryThreads * kTileQPL;
constexpr int kTileItemCap = 2048;                // records per tile list (shared memory: 16 KB as int2)
constexpr int kTileFrontCap = 1024;               // breadth-first frontier
constexpr int kTileFarCap = 512;                  // far set
constexpr int kTileSamples = 64;                  // 4^3 Chebyshev points
constexpr int kTileSampleStride = 72;             // 64 samples + centre(3) + 1/half-extent(3) + radius + pad
constexpr int kTileFallback = 1;                  // header flag: tile must be processed by the generic traversal

struct TileHeader
{
    int n_items;
    int n_far;
    int flags;
    int pad;
};

struct QueryArgs
{
    WnTreeView tree;
    float beta2;
    // points mode
    const float* q;       // [n*3]
    const unsigned* perm; // optional: slot -> point index (Morton order)
    int64_t n;
    int q_aligned16;
    // grid mode: block b covers lattice tile (b % tiles_x, (b / tiles_x) % tiles_y, b / (tiles_x*tiles_y) + tile_z0)
    GridDesc g;
    int tiles_x, tiles_y, tile_z0;
    // outputs (either may be null)
    float* out_omega;
    uint8_t* out_inside;
    unsigned long long* stats; // [4] tests, far-field evaluations, exact triangles, lane slots  (executed work)
    // tiled path
    int64_t tile_base;        // points mode: first tile of this launch
    TileHeader* plan_hdr;     // [tiles in launch]
    int2* plan_items;         // [tiles in launch][kTileItemCap]  (key, skip position)
    float* plan_samples;      // [tiles in launch][kTileSampleStride]
    float kappa;              // far set needs |c - P| >= kappa * tile radius
};

struct TravCounters
{
    unsigned long long T = 0, A = 0, E = 0, V = 0;
};

// Chebyshev points of the first kind, n = 4, and the Lagrange basis on them.
__device__ __forceinline__ float cheb_node(int k)
{
    return k == 0 ? 0.92387953251128674f : (k == 1 ? 0.38268343236508977f : (k == 2 ? -0.38268343236508977f : -0.92387953251128674f));
}
__device__ __forceinline__ void cheb_weights(float u, float w[4])
{
    const float a = 0.92387953251128674f, b = 0.38268343236508977f;
    const float den0 = (a - b) * (a + b) * (2.0f * a); // (x0-x1)(x0-x2)(x0-x3)
    const float den1 = (b - a) * (2.0f * b) * (b + a);
    const float d0 = u - a, d1 = u - b, d2 = u + b, d3 = u + a;
    w[0] = d1 * d2 * d3 * (1.0f / den0);
    w[1] = d0 * d2 * d3 * (1.0f / den1);
    w[2] = d0 * d1 * d3 * (-1.0f / den1);
    w[3] = d0 * d1 * d2 * (-1.0f / den0);
}

// list keys: (entry << 2) | (leaf << 1) | notest
__device__ __forceinline__ int tile_key(int entry, bool leaf, bool notest)
{
    return (entry << 2) | (leaf ? 2 : 0) | (notest ? 1 : 0);
}

// ----------------------------------------------------------------------------------------------------------------
// The shared traversal. LISTED = false: records are the packed tree itself. LISTED = true: records are the tile list in
// shared memory (key, skip position). Returns true if a far-field value that the list cannot recover from was not
// finite (the caller then redoes the tile generically; the reference descends in that case).
// ----------------------------------------------------------------------------------------------------------------
template <int QPL, bool STATS, bool LISTED>
__device__ __forceinline__ bool warp_traverse(const WnTreeView& t, const float beta2, const float (&qx)[QPL], const float (&qy)[QPL],
                                              const float (&qz)[QPL], const bool (&valid)[QPL], float (&acc)[QPL], const int2* s_items,
                                              const int n_items, TravCounters& cnt)
{
    const int lane = threadIdx.x & 31;
    const float4* __restrict__ r0 = t.rec[0];
    const float4* __restrict__ r1 = t.rec[1];
    const float4* __restrict__ r2 = t.rec[2];
    const float4* __restrict__ r3 = t.rec[3];
    const float4* __restrict__ r4 = t.rec[4];
    const float4* __restrict__ r5 = t.rec[5];
    const int* __restrict__ link = t.link;
    const float4* __restrict__ tris = t.tri;
    const int n = LISTED ? n_items : t.n_entries;
    int skip[QPL];
#pragma unroll
    for (int k = 0; k < QPL; ++k) {
        skip[k] = valid[k] ? 0 : n;
        acc[k] = 0.0f;
    }
    bool bad = false;
    // entry 0 is the root, which is never approximated (A.5): start at its first child unless the root is itself a leaf
    int i = LISTED ? 0 : (n > 1 ? 1 : 0);
    while (i < n) {
        int e = i, after = 0, lk = 0;
        bool leaf, notest = false;
        if (LISTED) {
            const int2 it = s_items[i];
            e = it.x >> 2;
            leaf = (it.x & 2) != 0;
            notest = (it.x & 1) != 0;
            after = it.y;
        }
        const float4 f0 = __ldg(r0 + e);
        if (!LISTED) {
            lk = __ldg(link + e);
            leaf = __float_as_int(f0.w) < 0;
            after = leaf ? i + 1 : lk;
        }
        // Unfused, like the reference: a decision flipped by an FMA's single rounding would change Omega by that record's
        // whole truncation error (~1e-4 * 4 pi at beta = 2).
        const float thr = __fmul_rn(fabsf(f0.w), beta2);
        float rx[QPL], ry[QPL], rz[QPL], l2[QPL];
        bool nearq[QPL], farq[QPL];
        bool anyfar = false;
#pragma unroll
        for (int k = 0; k < QPL; ++k) {
            const bool active = i >= skip[k];
            rx[k] = qx[k] - f0.x;
            ry[k] = qy[k] - f0.y;
            rz[k] = qz[k] - f0.z;
            l2[k] = __fadd_rn(__fadd_rn(__fmul_rn(rx[k], rx[k]), __fmul_rn(ry[k], ry[k])), __fmul_rn(rz[k], rz[k]));
            const bool nr = (LISTED && notest) ? false : (l2[k] <= thr);
            nearq[k] = active && nr;
            farq[k] = active && !nr;
            anyfar |= farq[k];
            if (STATS) cnt.T += (active && !(LISTED && notest)) ? 1 : 0;
        }
        if (STATS) cnt.V += (lane == 0) ? 32 * QPL : 0;
        if (__any_sync(kFull, anyfar)) {
            const float4 f1 = __ldg(r1 + e), f2 = __ldg(r2 + e), f3 = __ldg(r3 + e), f4 = __ldg(r4 + e), f5 = __ldg(r5 + e);
#pragma unroll
            for (int k = 0; k < QPL; ++k) {
                if (farq[k]) {
                    const float om = wn_eval_record(rx[k], ry[k], rz[k], l2[k], f1, f2, f3, f4, f5);
                    if (fabsf(om) <= 3.402823466e38f) {
                        acc[k] += om;
                        skip[k] = after;
                        if (STATS) ++cnt.A;
                    } else {
                        nearq[k] = true; // non-finite expansion: descend instead (A.5)
                        if (LISTED && notest) bad = true; // its children are not in the list
                    }
                }
            }
        }
        bool anynear = false;
#pragma unroll
        for (int k = 0; k < QPL; ++k) anynear |= nearq[k];
        anynear = __any_sync(kFull, anynear);
        if (leaf) {
            if (anynear) {
                if (LISTED) lk = __ldg(link + e);
                const int first = lk >> WN_LEAF_COUNT_BITS, count = (lk & (WN_MAX_LEAF_SIZE - 1)) + 1;
                for (int tt = 0; tt < count; ++tt) {
                    const float4 ta = __ldg(tris + 3 * (int64_t)(first + tt));
                    const float4 tb = __ldg(tris + 3 * (int64_t)(first + tt) + 1);
                    const float4 tc = __ldg(tris + 3 * (int64_t)(first + tt) + 2);
#pragma unroll
                    for (int k = 0; k < QPL; ++k) {
                        if (nearq[k]) {
                            acc[k] += wn_tri_solid_angle(qx[k], qy[k], qz[k], ta, tb, tc);
                            if (STATS) ++cnt.E;
                        }
                    }
                }
            }
            i = i + 1;
        } else {
            i = anynear ? i + 1 : after;
        }
    }
    return bad;
}

// ----------------------------------------------------------------------------------------------------------------
// Query point set-up shared by the kernels. Grid: CTA tile = 8 x 8 x (4*QPL) lattice points, warp tile 4 x 4 x (2*QPL).
// ----------------------------------------------------------------------------------------------------------------
template <int QPL>
__device__ __forceinline__ void grid_points(const QueryArgs& a, float (&qx)[QPL], float (&qy)[QPL], float (&qz)[QPL], bool (&valid)[QPL],
                                            int64_t (&oidx)[QPL])
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int bx = blockIdx.x % a.tiles_x;
    const int by = (blockIdx.x / a.tiles_x) % a.tiles_y;
    const int bz = blockIdx.x / (a.tiles_x * a.tiles_y) + a.tile_z0;
    const int x = bx * 8 + (wid & 1) * 4 + (lane & 3);
    const int y = by * 8 + ((wid >> 1) & 1) * 4 + ((lane >> 2) & 3);
#pragma unroll
    for (int k = 0; k < QPL; ++k) {
        const int z = a.g.z0 + bz * (4 * QPL) + (wid >> 2) * (2 * QPL) + 2 * k + (lane >> 4);
        valid[k] = x < a.g.nx && y < a.g.ny && z < a.g.z1;
        qx[k] = wn_lattice_coord(a.g.ox, a.g.sx, x);
        qy[k] = wn_lattice_coord(a.g.oy, a.g.sy, y);
        qz[k] = wn_lattice_coord(a.g.oz, a.g.sz, z);
        oidx[k] = valid[k] ? ((int64_t)(z - a.g.z0) * a.g.ny + y) * a.g.nx + x : -1;
    }
}

// Points: warp w of block b owns slots [(b*8 + w) * 32*QPL, +32*QPL); `stage` = 24*QPL float4 of shared memory per warp.
template <int QPL>
__device__ __forceinline__ void list_points(const QueryArgs& a, int64_t block, float4* stage, float (&qx)[QPL], float (&qy)[QPL],
                                            float (&qz)[QPL], bool (&valid)[QPL], int64_t (&oidx)[QPL])
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t wbase = (block * kQueryWarps + wid) * (32 * QPL);
    const bool staged = a.perm == nullptr && a.q_aligned16 && wbase + 32 * QPL <= a.n;
    if (staged) {
        // 32*QPL points = 96*QPL floats = 24*QPL float4, contiguous and 16-byte aligned: vectorised, coalesced
        const float4* src = reinterpret_cast<const float4*>(a.q + 3 * wbase);
        for (int j = lane; j < 24 * QPL; j += 32) stage[j] = __ldg(src + j);
        __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < QPL; ++k) {
        const int64_t s = wbase + k * 32 + lane;
        valid[k] = s < a.n;
        int64_t p = -1;
        qx[k] = qy[k] = qz[k] = 0.0f;
        if (valid[k]) {
            p = a.perm ? (int64_t)a.perm[s] : s;
            if (staged) {
                const float* f = reinterpret_cast<const float*>(stage) + 3 * (k * 32 + lane);
                qx[k] = f[0];
                qy[k] = f[1];
                qz[k] = f[2];
            } else {
                qx[k] = __ldg(a.q + 3 * p);
                qy[k] = __ldg(a.q + 3 * p + 1);
                qz[k] = __ldg(a.q + 3 * p + 2);
            }
        }
        oidx[k] = p;
    }
}

template <int QPL>
__device__ __forceinline__ void write_results(const QueryArgs& a, const int64_t (&oidx)[QPL], const float (&acc)[QPL])
{
#pragma unroll
    for (int k = 0; k < QPL; ++k) {
        if (oidx[k] >= 0) {
            if (a.out_omega) a.out_omega[oidx[k]] = acc[k];
            if (a.out_inside) a.out_inside[oidx[k]] = wn_inside_from_omega(acc[k]) ? 1 : 0;
        }
    }
}

__device__ __forceinline__ void flush_counters(const QueryArgs& a, TravCounters& c)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c.T += __shfl_xor_sync(kFull, c.T, o);
        c.A += __shfl_xor_sync(kFull, c.A, o);
        c.E += __shfl_xor_sync(kFull, c.E, o);
        c.V += __shfl_xor_sync(kFull, c.V, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(a.stats + 0, c.T);
        atomicAdd(a.stats + 1, c.A);
        atomicAdd(a.stats + 2, c.E);
        atomicAdd(a.stats + 3, c.V);
    }
}

// ---- generic kernel --------------------------------------------------------------------------------------------
template <int QPL, bool GRID, bool STATS>
__global__ void __launch_bounds__(kQueryThreads) k_query(const QueryArgs a)
{
    __shared__ float4 stage[GRID ? 1 : kQueryWarps * 24 * QPL];
    float qx[QPL], qy[QPL], qz[QPL], acc[QPL];
    bool valid[QPL];
    int64_t oidx[QPL];
    if (GRID)
        grid_points<QPL>(a, qx, qy, qz, valid, oidx);
    else
        list_points<QPL>(a, blockIdx.x, stage + (threadIdx.x >> 5) * 24 * QPL, qx, qy, qz, valid, oidx);
    TravCounters cnt;
    warp_traverse<QPL, STATS, false>(a.tree, a.beta2, qx, qy, qz, valid, acc, nullptr, 0, cnt);
    write_results<QPL>(a, oidx, acc);
    if (STATS) flush_counters(a, cnt);
}

// ---- tiled path: plan ------------------------------------------------------------------------------------------
// frontier words: entry | (has_mixed_ancestor << 30)
__device__ __forceinline__ void tile_push_children(const WnTreeView& t, int e, int flag, int* front, int* count, int* overflow)
{
    const int end = __ldg(t.link + e);
    int c = e + 1;
    while (c < end) {
        const int pos = atomicAdd(count, 1);
        if (pos < kTileFrontCap)
            front[pos] = c | (flag << 30);
        else
            *overflow = 1;
        const bool leaf = __float_as_int(__ldg(&t.rec[0][c].w)) < 0;
        c = leaf ? c + 1 : __ldg(t.link + c);
    }
}

template <bool GRID>
__global__ void __launch_bounds__(kQueryThreads) k_tile_plan(const QueryArgs a)
{
    __shared__ int s_front[2][kTileFrontCap];
    __shared__ int s_items[kTileItemCap];
    __shared__ int s_far[kTileFarCap];
    __shared__ float s_samp[kQueryWarps][kTileSamples];
    __shared__ int s_cnt[8]; // 0,1 frontier sizes; 2 items; 3 far; 4 overflow / bad
    __shared__ float s_geo[8];
    __shared__ float s_red[kQueryWarps][6];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const WnTreeView& t = a.tree;

    // ---- tile bounding sphere ------------------------------------------------------------------------------------
    if (GRID) {
        if (tid == 0) {
            const int bx = blockIdx.x % a.tiles_x;
            const int by = (blockIdx.x / a.tiles_x) % a.tiles_y;
            const int bz = blockIdx.x / (a.tiles_x * a.tiles_y) + a.tile_z0;
            const float lx = wn_lattice_coord(a.g.ox, a.g.sx, bx * 8), hx = wn_lattice_coord(a.g.ox, a.g.sx, bx * 8 + 7);
            const float ly = wn_lattice_coord(a.g.oy, a.g.sy, by * 8), hy = wn_lattice_coord(a.g.oy, a.g.sy, by * 8 + 7);
            const float lz = wn_lattice_coord(a.g.oz, a.g.sz, a.g.z0 + bz * 8), hz = wn_lattice_coord(a.g.oz, a.g.sz, a.g.z0 + bz * 8 + 7);
            s_red[0][0] = fminf(lx, hx);
            s_red[0][1] = fminf(ly, hy);
            s_red[0][2] = fminf(lz, hz);
            s_red[0][3] = fmaxf(lx, hx);
            s_red[0][4] = fmaxf(ly, hy);
            s_red[0][5] = fmaxf(lz, hz);
        }
    } else {
        float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
        const int64_t base = ((int64_t)blockIdx.x + a.tile_base) * kTileQueries;
#pragma unroll
        for (int k = 0; k < kTileQPL; ++k) {
            const int64_t s = base + k * kQueryThreads + tid;
            if (s < a.n) {
                const int64_t p = a.perm ? (int64_t)a.perm[s] : s;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float c = __ldg(a.q + 3 * p + d);
                    lo[d] = fminf(lo[d], c); // NaN coordinates are ignored by fminf/fmaxf
                    hi[d] = fmaxf(hi[d], c);
                }
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo[d] = fminf(lo[d], __shfl_xor_sync(kFull, lo[d], o));
                hi[d] = fmaxf(hi[d], __shfl_xor_sync(kFull, hi[d], o));
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                s_red[wid][d] = lo[d];
                s_red[wid][3 + d] = hi[d];
            }
        }
    }
    if (tid < 8) s_cnt[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        float lo[3], hi[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            lo[d] = s_red[0][d];
            hi[d] = s_red[0][3 + d];
            if (!GRID) {
                for (int w = 1; w < kQueryWarps; ++w) {
                    lo[d] = fminf(lo[d], s_red[w][d]);
                    hi[d] = fmaxf(hi[d], s_red[w][3 + d]);
                }
            }
        }
        float r2 = 0.0f;
        bool finite = true;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float c = 0.5f * (lo[d] + hi[d]), h = 0.5f * (hi[d] - lo[d]);
            s_geo[d] = c;
            s_geo[3 + d] = h > 0.0f ? 1.0f / h : 0.0f;
            r2 += h * h;
            finite = finite && (fabsf(c) <= 3.0e38f) && (h >= 0.0f) && (h <= 3.0e38f);
        }
        // inflated so that rounding in the distance computations below can never misclassify a point of the tile
        s_geo[6] = sqrtf(r2) * 1.0001f + 1e-30f;
        s_geo[7] = 0.0f;
        if (!finite) s_cnt[4] = 1; // non-finite coordinates (or an empty tile): generic path
        const int n_entries = t.n_entries;
        if (n_entries > 1) {
            tile_push_children(t, 0, 0, s_front[0], &s_cnt[0], &s_cnt[4]);
        } else if (n_entries == 1) {
            s_items[0] = tile_key(0, true, false);
            s_cnt[2] = 1;
        }
    }
    __syncthreads();
    const float cx = s_geo[0], cy = s_geo[1], cz = s_geo[2], ra = s_geo[6];
    const float hx = s_geo[3] > 0.0f ? 1.0f / s_geo[3] : 0.0f, hy = s_geo[4] > 0.0f ? 1.0f / s_geo[4] : 0.0f,
                hz = s_geo[5] > 0.0f ? 1.0f / s_geo[5] : 0.0f;

    // ---- breadth-first classification ----------------------------------------------------------------------------
    int cur = 0;
    while (true) {
        const int F = min(s_cnt[cur], kTileFrontCap);
        if (F == 0 || s_cnt[4]) break;
        int* nxt = s_front[cur ^ 1];
        for (int idx = tid; idx < F; idx += kQueryThreads) {
            const int word = s_front[cur][idx];
            const int e = word & 0x3fffffff, manc = (word >> 30) & 1;
            const float4 f0 = __ldg(t.rec[0] + e);
            const bool leaf = __float_as_int(f0.w) < 0;
            const float thr = fabsf(f0.w) * a.beta2;
            const float dx = cx - f0.x, dy = cy - f0.y, dz = cz - f0.z;
            const float D = sqrtf(dx * dx + dy * dy + dz * dz);
            const float dm = D - ra, dp = D + ra;
            const bool allfar = dm > 0.0f && dm * dm > thr * 1.0001f;
            const bool allnear = dp * dp <= thr * 0.9999f;
            if (allfar) {
                // far set: the record's field must be smooth across the tile, i.e. the tile is small against its distance
                // both to the expansion centre and to the nearest possible source point (bounding sphere of radius R)
                if (!manc && D >= a.kappa * ra && D - sqrtf(fabsf(f0.w)) >= 0.5f * a.kappa * ra) {
                    const int pos = atomicAdd(&s_cnt[3], 1);
                    if (pos < kTileFarCap)
                        s_far[pos] = e;
                    else
                        s_cnt[4] = 1;
                } else {
                    const int pos = atomicAdd(&s_cnt[2], 1);
                    if (pos < kTileItemCap)
                        s_items[pos] = tile_key(e, leaf, true);
                    else
                        s_cnt[4] = 1;
                }
            } else if (allnear && !leaf) {
                tile_push_children(t, e, manc, nxt, &s_cnt[cur ^ 1], &s_cnt[4]);
            } else {
                const int pos = atomicAdd(&s_cnt[2], 1);
                if (pos < kTileItemCap)
                    s_items[pos] = tile_key(e, leaf, false);
                else
                    s_cnt[4] = 1;
                if (!leaf && !allnear) tile_push_children(t, e, 1, nxt, &s_cnt[cur ^ 1], &s_cnt[4]);
            }
        }
        __syncthreads();
        if (tid == 0) s_cnt[cur] = 0;
        cur ^= 1;
        __syncthreads();
    }
    __syncthreads();
    const int n_items = min(s_cnt[2], kTileItemCap), n_far = min(s_cnt[3], kTileFarCap);
    bool fallback = s_cnt[4] != 0;

    // ---- back to depth-first order: bitonic sort of the keys, then skip links in list coordinates -------------------
    if (!fallback) {
        int N = 2;
        while (N < n_items) N <<= 1;
        for (int j = n_items + tid; j < N; j += kQueryThreads) s_items[j] = 0x7fffffff;
        __syncthreads();
        for (int k = 2; k <= N; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < N; i += kQueryThreads) {
                    const int p = i ^ j;
                    if (p > i) {
                        const int va = s_items[i], vb = s_items[p];
                        const bool asc = (i & k) == 0;
                        if ((va > vb) == asc) {
                            s_items[i] = vb;
                            s_items[p] = va;
                        }
                    }
                }
                __syncthreads();
            }
        }
        int2* out = a.plan_items + (int64_t)blockIdx.x * kTileItemCap;
        for (int j = tid; j < n_items; j += kQueryThreads) {
            const int key = s_items[j];
            const int e = key >> 2;
            const int end = (key & 2) ? e + 1 : __ldg(t.link + e);
            const int target = end << 2;
            int lo = j + 1, hi = n_items; // first position whose key >= target
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_items[mid] < target)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            out[j] = make_int2(key, lo);
        }
    }

    // ---- far set: sample its field at the 4^3 Chebyshev points of the tile's box -----------------------------------
    float sacc[2] = {0.0f, 0.0f};
    bool bad = false;
    if (!fallback) {
        float px[2], py[2], pz[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int s = lane + 32 * k;
            px[k] = cx + hx * cheb_node(s & 3);
            py[k] = cy + hy * cheb_node((s >> 2) & 3);
            pz[k] = cz + hz * cheb_node(s >> 4);
        }
        for (int m = wid; m < n_far; m += kQueryWarps) {
            const int e = s_far[m];
            const float4 f0 = __ldg(t.rec[0] + e), f1 = __ldg(t.rec[1] + e), f2 = __ldg(t.rec[2] + e), f3 = __ldg(t.rec[3] + e),
                         f4 = __ldg(t.rec[4] + e), f5 = __ldg(t.rec[5] + e);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float rx = px[k] - f0.x, ry = py[k] - f0.y, rz = pz[k] - f0.z;
                const float l2 = rx * rx + ry * ry + rz * rz;
                const float om = wn_eval_record(rx, ry, rz, l2, f1, f2, f3, f4, f5);
                bad = bad || !(fabsf(om) <= 3.402823466e38f);
                sacc[k] += om;
            }
        }
        s_samp[wid][lane] = sacc[0];
        s_samp[wid][lane + 32] = sacc[1];
    }
    fallback = __syncthreads_or((int)(bad || fallback)) != 0;
    float* sout = a.plan_samples + (int64_t)blockIdx.x * kTileSampleStride;
    if (!fallback && tid < kTileSamples) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < kQueryWarps; ++w) s += s_samp[w][tid];
        sout[tid] = s;
    }
    if (tid < 8) sout[kTileSamples + tid] = s_geo[tid];
    if (tid == 0) {
        TileHeader h;
        h.n_items = fallback ? 0 : n_items;
        h.n_far = fallback ? 0 : n_far;
        h.flags = fallback ? kTileFallback : 0;
        h.pad = 0;
        a.plan_hdr[blockIdx.x] = h;
        if (a.stats) {
            // executed work of the plan: far-set evaluations at the sample points (counted as far-field evaluations)
            atomicAdd(a.stats + 1, (unsigned long long)(fallback ? 0 : n_far) * kTileSamples);
        }
    }
}

// ---- tiled path: query -----------------------------------------------------------------------------------------
template <bool GRID, bool STATS>
__global__ void __launch_bounds__(kQueryThreads) k_tile_query(const QueryArgs a)
{
    __shared__ int2 s_items[kTileItemCap];
    __shared__ float s_samp[kTileSampleStride];
    __shared__ float4 stage[GRID ? 1 : kQueryWarps * 24 * kTileQPL];
    constexpr int QPL = kTileQPL;
    float qx[QPL], qy[QPL], qz[QPL], acc[QPL];
    bool valid[QPL];
    int64_t oidx[QPL];
    if (GRID)
        grid_points<QPL>(a, qx, qy, qz, valid, oidx);
    else
        list_points<QPL>(a, (int64_t)blockIdx.x + a.tile_base, stage + (threadIdx.x >> 5) * 24 * QPL, qx, qy, qz, valid, oidx);
    const TileHeader hdr = a.plan_hdr[blockIdx.x];
    const bool fallback = (hdr.flags & kTileFallback) != 0;
    if (!fallback) {
        const int2* src = a.plan_items + (int64_t)blockIdx.x * kTileItemCap;
        for (int j = threadIdx.x; j < hdr.n_items; j += kQueryThreads) s_items[j] = src[j];
        if (threadIdx.x < kTileSampleStride) s_samp[threadIdx.x] = a.plan_samples[(int64_t)blockIdx.x * kTileSampleStride + threadIdx.x];
    }
    __syncthreads();
    TravCounters cnt;
    bool bad = false;
    if (!fallback) bad = warp_traverse<QPL, STATS, true>(a.tree, a.beta2, qx, qy, qz, valid, acc, s_items, hdr.n_items, cnt);
    if (__syncthreads_or((int)(bad || fallback))) {
        // generic traversal of the whole tree for this tile (list overflow, non-finite far field, degenerate tile)
        warp_traverse<QPL, STATS, false>(a.tree, a.beta2, qx, qy, qz, valid, acc, nullptr, 0, cnt);
    } else {
        const float cx = s_samp[64], cy = s_samp[65], cz = s_samp[66];
        const float ihx = s_samp[67], ihy = s_samp[68], ihz = s_samp[69];
#pragma unroll
        for (int k = 0; k < QPL; ++k) {
            float wx[4], wy[4], wz[4];
            cheb_weights((qx[k] - cx) * ihx, wx);
            cheb_weights((qy[k] - cy) * ihy, wy);
            cheb_weights((qz[k] - cz) * ihz, wz);
            float far = 0.0f;
#pragma unroll
            for (int kz = 0; kz < 4; ++kz) {
                float sz = 0.0f;
#pragma unroll
                for (int ky = 0; ky < 4; ++ky) {
                    const float4 row = *reinterpret_cast<const float4*>(&s_samp[(kz * 4 + ky) * 4]);
                    sz += wy[ky] * (wx[0] * row.x + wx[1] * row.y + wx[2] * row.z + wx[3] * row.w);
                }
                far += wz[kz] * sz;
            }
            acc[k] += far;
        }
    }
    write_results<QPL>(a, oidx, acc);
    if (STATS) flush_counters(a, cnt);
}

} // namespace wn
