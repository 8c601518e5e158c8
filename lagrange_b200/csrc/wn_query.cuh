// wn_query.cuh — sm_100a tree query kernels (K6). Replace UT_SolidAngle::computeSolidAngle
// (modules/winding/src/FastWindingNumber.cpp:66,75) for whole batches.
//
// warp_traverse  : the per-point traversal. One warp owns 32*QPL spatially adjacent queries and walks a depth-first
//                  sequence of records once for all of them (stackless: "descend" = next record, "skip subtree" = the
//                  record's skip link). Each lane keeps the reference's per-point semantics through a private resume
//                  index: a lane that accepted a far-field record ignores records until the end of that subtree, lanes
//                  that must descend keep going; the warp leaves a subtree only when no lane needs it. The accept test
//                  |q-P|^2 <= beta^2 R^2 is formed unfused, exactly like the reference, so every point takes the same
//                  branches as in the CPU algorithm. Far field = folded order-2 Taylor record, leaves = exact
//                  Van Oosterom-Strackee triangles. Record reads are warp-uniform float4 broadcasts.
// k_query        : generic kernel: the record sequence is the whole packed tree (small or incoherent batches, fallback).
// k_tile_plan +  : tiled path for coherent batches (lattices, Morton-sorted point sets). A tile is 8x8x8 lattice points
// k_tile_query     (or 512 consecutive sorted points). k_tile_plan classifies tree nodes against the tile's bounding
//                  sphere, breadth first, and sorts the survivors back into depth-first order:
//                    far set   : far for every point, no ancestor that some point accepted, and smooth over the tile ->
//                                evaluated ONCE PER TILE at 4^3 Chebyshev points, interpolated per query
//                    direct    : far for every point, no such ancestor, but too close to interpolate -> every point
//                                evaluates it, nobody tests it; records gathered contiguously for the query kernel
//                    exact     : leaf that is near for every point, no such ancestor -> every point evaluates its
//                                triangles exactly, nobody tests it; triangles gathered contiguously
//                    (near for every point, internal -> dropped, children expanded)
//                    conditional: everything under a node that only some points accept -> short list with skip links in
//                                list coordinates, walked by warp_traverse with the full per-point logic
//                  k_tile_query streams the gathered records/triangles through tight loops (no tests, no votes, no
//                  divergence), runs warp_traverse over the conditional list (read through L1, warp-uniform) and adds the
//                  tensor-product Chebyshev interpolant of the far set. Which records a point accepts is unchanged; only
//                  where their sum is evaluated differs (interpolation error ~1e-5 * 4 pi, bounded in tests).
// All FP32 CUDA-core work (FMA pipe + MUFU rsqrt/atan): no tensor cores by design (BASELINE.json north_star).
#pragma once

#include <cuda_runtime.h>

#include "wn_exact.cuh"

namespace wn {

// resident CTAs per SM the query kernels are compiled for (register budget = 65536 / (256 * N)). Measured on cfg2:
// 4 -> 4.49, 5 -> 4.63 (48 registers, 8 B spilled), 6 -> 4.36 G queries/s.
#ifndef WN_Q_MIN_CTAS
#define WN_Q_MIN_CTAS 5
#endif
#ifndef WN_TQ_MIN_CTAS
#define WN_TQ_MIN_CTAS 5
#endif
constexpr int kQueryThreads = 256;
constexpr int kQueryWarps = kQueryThreads / 32;
#ifndef WN_TILE_QPL
#define WN_TILE_QPL 2
#endif
constexpr int kTileQPL = WN_TILE_QPL;             // queries per lane in k_tile_query: 2 (8 warp tasks of 4x4x4 points per tile) or 4 (4 tasks of 4x4x8)
constexpr int kTileQueries = 512;                 // a tile is 8^3 lattice points / 512 consecutive sorted points whatever the split
constexpr int kTileTasks = kTileQueries / (32 * kTileQPL);
#ifndef WN_PLAN_THREADS
#define WN_PLAN_THREADS 128
#endif
constexpr int kPlanThreads = WN_PLAN_THREADS;                 // k_tile_plan: latency bound, small CTAs so that many are resident
constexpr int kPlanWarps = kPlanThreads / 32;
#ifndef WN_TILE_CAP_SCALE
#define WN_TILE_CAP_SCALE 1
#endif
// k_tile_plan is latency-bound (barrier per breadth-first round), so resident CTAs count: list capacities of 3/4 of round 1's
// (1536 / 768 / 768: 21 KB of shared memory) and 56 registers put 9 CTAs on an SM instead of 8. Measured on cfg2 (G q/s), capacity
// quarters x CTAs/SM: 4 x 8 (64 regs) 7.92; 3 x 9 (56 regs) 8.03; 3 x 10 (48 regs, 44 B spilled) 7.97; 2 x 12 (40 regs, 100 B
// spilled) 7.99; uncapped registers (112, 4 CTAs) 7.27. cfg3 +1 %, forced-tiled cfg5 +4 %, overflow fallbacks unchanged.
#ifndef WN_TILE_CAP_Q
#define WN_TILE_CAP_Q 3
#endif
#ifndef WN_PLAN_MIN_CTAS
#define WN_PLAN_MIN_CTAS 9
#endif
#define WN_PLAN_BOUNDS __launch_bounds__(kPlanThreads, WN_PLAN_MIN_CTAS)
constexpr int kTileAllCap = 512 * WN_TILE_CAP_Q * WN_TILE_CAP_SCALE;   // classified records per tile (conditional + direct + exact)
constexpr int kTileFrontCap = 256 * WN_TILE_CAP_Q * WN_TILE_CAP_SCALE; // breadth-first frontier
constexpr int kTileFarCap = 512;                  // far set
constexpr int kTileDirCap = 512;                  // direct records
constexpr int kTileExactCap = 256 * WN_TILE_CAP_Q * WN_TILE_CAP_SCALE; // exact leaves (also bounded by the frontier buffer reused for offsets)
constexpr int kPlanMaxRounds = 96;                // breadth-first rounds = hierarchy depth bound (deeper: generic path)
constexpr int kTileSamples = 64;                  // 4^3 Chebyshev points
constexpr int kTileSampleStride = 72;             // 64 samples + centre(3) + 1/half-extent(3) + radius + pad
constexpr int kTileFallback = 1;                  // header flag: tile must be processed by the generic traversal

// record classes of the tile plan (low bits of a key = (entry << 3) | (leaf << 2) | class)
constexpr int kClsCond = 0;       // tested per point
constexpr int kClsCondFar = 1;    // far for every point of the tile, but under a node only some points accept: no test
constexpr int kClsDirect = 2;     // far for every point, unconditional
constexpr int kClsExact = 3;      // leaf, near for every point, unconditional

struct TileHeader
{
    int n_cond, n_dir, n_tri, flags;
    long long offset; // byte offset of the tile's packet in the arena: [cond 32 B x n_cond | dir 96 B x n_dir | tri 48 B x n_tri]
    long long pad;
};

// Entry list of a planning block above the tiles (hierarchical planning, see k_plan_block): the hierarchy nodes its child
// blocks / tiles have to classify themselves, as ints in the arena.
struct PlanBlockHeader
{
    long long offset; // byte offset of the list in the arena
    int count;
    int flags;        // kTileFallback: the block overflowed; its children plan from the root and take no samples from it
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
    int layer_step, out_layer0; // strided layers: CTA layer l -> lattice layer tile_z0 + l*layer_step, output layer out_layer0 + l
    // diagonal sharding (wn_query_grid_sharded): the lattice is cut in shard_q parts along y (part_ny rows each, a multiple of 8;
    // tiles_y counts the tile rows of ONE part) and a layer of this call covers one part only: layer lz -> part
    // ((shard_r - lz) / shard_c) mod shard_q. shard_q <= 1: whole rows (part_ny = ny).
    int shard_q, shard_c, shard_r, part_ny;
    // outputs (either may be null)
    float* out_omega;
    uint8_t* out_inside;
    unsigned long long* stats; // [4] tests, far-field evaluations, exact triangles, lane slots  (executed work)
    // tiled path
    int64_t tile_base;        // points mode: first tile of this launch
    TileHeader* plan_hdr;     // [tiles in launch]
    float* plan_samples;      // [tiles in launch][kTileSampleStride]
    char* plan_arena;         // variable-size packets
    unsigned long long* plan_cursor; // bump allocator over the arena (reset before every k_tile_plan launch)
    long long plan_arena_bytes;
    float kappa;              // far set needs |c - P| >= kappa * tile radius and |c - P| - R >= kappa/2 * tile radius
    int tiles_per_cta, launch_tiles; // k_tile_query: consecutive tiles per CTA, tiles in this launch
    int run_ctas;             // k_tile_query: the first run_ctas CTAs take tiles_per_cta tiles each, the CTAs after them one tile each
    int* tile_order;          // [launch_tiles] tiles with long conditional lists first (written by k_tile_plan, counters after plan_cursor)
    int heavy_cond;           // a tile is 'heavy' from this many conditional records on
    int probe_stride;         // > 0: k_tile_plan only classifies every probe_stride-th tile and adds the class sizes to `probe`
    unsigned long long* probe; // [5] far, conditional, direct, exact records, fallback tiles
    unsigned long long* trace; // diagnostics (WN_TRACE_FILE): per CTA of this launch {start ns, end ns, SM id, 0}; null in production
    // hierarchical planning (lattices): a block of level k covers 2^k x 2^k x (2^k | 1) tiles. The level being planned reads its
    // parent level (up_*), the block kernel writes lvl_*.
    const PlanBlockHeader* up_hdr; // parent level's entry lists; null: plan from the root
    const float* up_samples;       // parent level's far-field samples [blocks][kTileSampleStride]
    int up_bx, up_by, up_zs;       // parent level: blocks per row / column, 1 if the parent halves the layer index too, else 0
    PlanBlockHeader* lvl_hdr;      // k_plan_block output
    float* lvl_samples;
    int lvl_shift, lvl_zshift, lvl_bx, lvl_by;
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

__device__ __forceinline__ int tile_key(int entry, bool leaf, int cls)
{
    return (entry << 3) | (leaf ? 4 : 0) | cls;
}

// accessors of the hot/cold record layout (wn_device.cuh, WnTreeView)
__device__ __forceinline__ float4 rec_hot(const WnTreeView& t, int e, int k)
{
    return __ldg(t.hot + 2 * (int64_t)e + k);
}
__device__ __forceinline__ float4 rec_cold(const WnTreeView& t, int e, int k)
{
    return __ldg(t.cold + 4 * (int64_t)e + k);
}
__device__ __forceinline__ int rec_link(const WnTreeView& t, int e)
{
    return __float_as_int(__ldg(&t.hot[2 * (int64_t)e + 1].w));
}

// wn_eval_record (wn_device.cuh) with its same-shaped chains advanced two at a time by sm_100's packed FP32 instructions
// (FMUL2 / FFMA2: one issue slot, two correctly rounded results, bit-identical to the scalar sequence). The record's storage
// order already is the operand-pair order, so the pairs come straight out of the 128-bit loads; only the unit vector has to
// be duplicated (3 MOV). 30 issue slots instead of 39 — these kernels are issue-bound, not FMA-pipe-bound.
__device__ __forceinline__ float eval_record(float rx, float ry, float rz, float l2, const float4& f1, const float4& c0, const float4& c1,
                                             const float4& c2, const float4& c3)
{
    const float m1 = wn_rsqrt_ftz(l2);
    const float x = __fmul_rn(rx, m1), y = __fmul_rn(ry, m1), z = __fmul_rn(rz, m1);
    const float m2 = __fmul_rn(m1, m1);
    const float2 X = make_float2(x, x), Y = make_float2(y, y), Z = make_float2(z, z);
    const float2 tp = __ffma2_rn(Z, make_float2(c1.x, c1.y), __ffma2_rn(Y, make_float2(c0.z, c0.w), __fmul2_rn(X, make_float2(c0.x, c0.y))));
    const float2 uq = __ffma2_rn(Z, make_float2(c2.x, c2.y), __fmul2_rn(Y, make_float2(c1.z, c1.w)));
    const float2 ws = __fmul2_rn(Z, make_float2(c2.z, c2.w));
    const float2 a1cx = __ffma2_rn(Z, ws, __ffma2_rn(Y, uq, __fmul2_rn(X, tp)));
    const float2 ng = __ffma2_rn(Z, make_float2(c3.z, c3.w), __fmul2_rn(Y, make_float2(c3.x, c3.y)));
    const float n = __fmaf_rn(x, f1.z, ng.x);
    const float2 hk = __fmul2_rn(Z, make_float2(f1.x, f1.y));
    const float cy = __fmaf_rn(z, hk.x, __fmul_rn(y, ng.y));
    const float cz = __fmul_rn(z, hk.y);
    const float a2 = __fmaf_rn(z, cz, __fmaf_rn(y, cy, __fmul_rn(x, a1cx.y)));
    return __fmul_rn(m2, __fmaf_rn(m1, __fmaf_rn(m1, a2, a1cx.x), -n));
}

// ----------------------------------------------------------------------------------------------------------------
// The per-point traversal. LISTED = false: records are the packed tree itself. LISTED = true: records are the tile's
// conditional list in the plan's packet (key, skip position). Accumulates into acc. Returns true if a far-field value the
// list cannot recover from was not finite (the caller then redoes the tile generically; the reference descends in
// that case).
// ----------------------------------------------------------------------------------------------------------------
template <int QPL, bool STATS, bool LISTED>
__device__ __forceinline__ bool warp_traverse(const WnTreeView& t, const float beta2, const float (&qx)[QPL], const float (&qy)[QPL],
                                              const float (&qz)[QPL], const bool (&valid)[QPL], float (&acc)[QPL], const int2* s_items,
                                              const int n_items, TravCounters& cnt)
{
    const int lane = threadIdx.x & 31;
    const float4* __restrict__ hot = t.hot;
    const float4* __restrict__ cold = t.cold;
    const float4* __restrict__ tris = t.tri;
    const int n = LISTED ? n_items : t.n_entries;
    int skip[QPL];
#pragma unroll
    for (int k = 0; k < QPL; ++k) skip[k] = valid[k] ? 0 : n;
    bool bad = false;
    // entry 0 is the root, which is never approximated (A.5): start at its first child unless the root is itself a leaf
    int i = LISTED ? 0 : (n > 1 ? 1 : 0);
    while (i < n) {
        int e = i, after = 0;
        bool leaf, notest = false;
        if (LISTED) {
            const int2 it = __ldg(s_items + i);
            e = it.x >> 3;
            leaf = (it.x & 4) != 0;
            notest = (it.x & 3) == kClsCondFar;
            after = it.y;
        }
        // one address, one 32-byte sector: centre + radius, normal + link
        const float4* __restrict__ hp = hot + 2 * (int64_t)e;
        const float4 f0 = __ldg(hp), f1 = __ldg(hp + 1);
        const int lk = __float_as_int(f1.w);
        if (!LISTED) {
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
            const float4* __restrict__ cp = cold + 4 * (int64_t)e;
            const float4 f2 = __ldg(cp), f3 = __ldg(cp + 1), f4 = __ldg(cp + 2), f5 = __ldg(cp + 3);
#pragma unroll
            for (int k = 0; k < QPL; ++k) {
                if (farq[k]) {
                    const float om = eval_record(rx[k], ry[k], rz[k], l2[k], f1, f2, f3, f4, f5);
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
// The per-point traversal over a tile's conditional list (k_tile_query). Same decisions as warp_traverse, but the list is
// self-contained for the test: an item is 32 bytes, (Px, Py, Pz, beta^2 R^2) + (key, skip position, leaf link, -), written by the
// plan, so a step is two consecutive 128-bit loads instead of list -> record (a dependent pair) and the tree is touched only
// when some lane actually evaluates. A non-finite far-field value makes the warp redo its sub-block with the generic traversal
// (the reference descends in that case; that path handles it point by point).
// ----------------------------------------------------------------------------------------------------------------
template <int QPL, bool STATS>
__device__ __forceinline__ bool tile_cond_walk(const WnTreeView& t, const float (&qx)[QPL], const float (&qy)[QPL], const float (&qz)[QPL],
                                               const bool (&valid)[QPL], float (&acc)[QPL], const float4* __restrict__ items, const int n_items,
                                               TravCounters& cnt)
{
    const int lane = threadIdx.x & 31;
    const float4* __restrict__ hot = t.hot;
    const float4* __restrict__ cold = t.cold;
    const float4* __restrict__ tris = t.tri;
    int skip[QPL];
#pragma unroll
    for (int k = 0; k < QPL; ++k) skip[k] = valid[k] ? 0 : n_items;
    bool bad = false;
    int i = 0;
    while (i < n_items) {
        const float4 c0 = __ldg(items + 2 * i);
        const int4 c1 = __ldg(reinterpret_cast<const int4*>(items + 2 * i + 1));
        const int e = c1.x >> 3, after = c1.y;
        const bool leaf = (c1.x & 4) != 0, notest = (c1.x & 3) == kClsCondFar;
        float rx[QPL], ry[QPL], rz[QPL], l2[QPL];
        bool nearq[QPL], farq[QPL];
        bool anyfar = false, anynear = false;
#pragma unroll
        for (int k = 0; k < QPL; ++k) {
            const bool active = i >= skip[k];
            rx[k] = qx[k] - c0.x;
            ry[k] = qy[k] - c0.y;
            rz[k] = qz[k] - c0.z;
            // unfused, like the reference (see warp_traverse)
            l2[k] = __fadd_rn(__fadd_rn(__fmul_rn(rx[k], rx[k]), __fmul_rn(ry[k], ry[k])), __fmul_rn(rz[k], rz[k]));
            const bool nr = !notest && l2[k] <= c0.w;
            nearq[k] = active && nr;
            farq[k] = active && !nr;
            anyfar |= farq[k];
            anynear |= nearq[k];
            if (STATS) cnt.T += (active && !notest) ? 1 : 0;
        }
        if (STATS) cnt.V += (lane == 0) ? 32 * QPL : 0;
        if (__any_sync(kFull, anyfar)) {
            const float4* __restrict__ cp = cold + 4 * (int64_t)e;
            const float4 f1 = __ldg(hot + 2 * (int64_t)e + 1);
            const float4 f2 = __ldg(cp), f3 = __ldg(cp + 1), f4 = __ldg(cp + 2), f5 = __ldg(cp + 3);
#pragma unroll
            for (int k = 0; k < QPL; ++k) {
                if (farq[k]) {
                    // (checked per evaluation: checking acc once after the walk allocates registers worse under the 48-register
                    // cap -- 40 B of spills in the loop, measured 2.5 % slower)
                    const float om = eval_record(rx[k], ry[k], rz[k], l2[k], f1, f2, f3, f4, f5);
                    bad = bad || !(fabsf(om) <= 3.402823466e38f);
                    acc[k] += om;
                    skip[k] = after;
                    if (STATS) ++cnt.A;
                }
            }
        }
        anynear = __any_sync(kFull, anynear);
        if (leaf) {
            if (anynear) {
                const int lk = c1.z;
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
// Lattice tile layer and first row of CTA layer l of this launch (see QueryArgs: strided layers, diagonal sharding).
__device__ __forceinline__ void layer_geom(const QueryArgs& a, const int l, int& bz, int& y0)
{
    bz = a.tile_z0 + l * a.layer_step;
    y0 = 0;
    if (a.shard_q > 1) {
        int q = ((a.shard_r - bz) / a.shard_c) % a.shard_q; // shard_r - bz is a multiple of shard_c by construction
        if (q < 0) q += a.shard_q;
        y0 = q * a.part_ny;
    }
}

template <int QPL>
__device__ __forceinline__ void grid_points(const QueryArgs& a, const int block, const int wid, float (&qx)[QPL], float (&qy)[QPL],
                                            float (&qz)[QPL], bool (&valid)[QPL], int64_t (&oidx)[QPL])
{
    const int lane = threadIdx.x & 31;
    const int bx = block % a.tiles_x;
    const int by = (block / a.tiles_x) % a.tiles_y;
    // CTA layer l of this launch covers the lattice layer tile_z0 + l * layer_step (strided sharding across GPUs: every
    // layer_step-th layer, results stored compactly: local layer out_layer0 + l)
    const int l = block / (a.tiles_x * a.tiles_y);
    int bz, y0;
    layer_geom(a, l, bz, y0);
    const int x = bx * 8 + (wid & 1) * 4 + (lane & 3);
    const int yl = by * 8 + ((wid >> 1) & 1) * 4 + ((lane >> 2) & 3); // row inside the part (= the lattice row when not sharded in y)
    const int y = y0 + yl;
#pragma unroll
    for (int k = 0; k < QPL; ++k) {
        const int dz = (wid >> 2) * (2 * QPL) + 2 * k + (lane >> 4);
        const int z = a.g.z0 + bz * (4 * QPL) + dz;
        const int zl = (a.out_layer0 + l) * (4 * QPL) + dz;
        valid[k] = x < a.g.nx && y < a.g.ny && z < a.g.z1;
        qx[k] = wn_lattice_coord(a.g.ox, a.g.sx, x);
        qy[k] = wn_lattice_coord(a.g.oy, a.g.sy, y);
        qz[k] = wn_lattice_coord(a.g.oz, a.g.sz, z);
        oidx[k] = valid[k] ? ((int64_t)zl * a.part_ny + yl) * a.g.nx + x : -1;
    }
}

// k_tile_query with 4 queries per lane: task `sub` (0..3) of an 8x8x8 tile = the 4x4x8 column (sub & 1, sub >> 1); lane -> (x, y,
// z parity), query k -> z pair k
template <int QPL>
__device__ __forceinline__ void tile_column_points(const QueryArgs& a, const int block, const int sub, float (&qx)[QPL], float (&qy)[QPL],
                                                   float (&qz)[QPL], bool (&valid)[QPL], int64_t (&oidx)[QPL])
{
    const int lane = threadIdx.x & 31;
    const int bx = block % a.tiles_x;
    const int by = (block / a.tiles_x) % a.tiles_y;
    const int l = block / (a.tiles_x * a.tiles_y);
    int bz, y0;
    layer_geom(a, l, bz, y0);
    const int x = bx * 8 + (sub & 1) * 4 + (lane & 3);
    const int yl = by * 8 + (sub >> 1) * 4 + ((lane >> 2) & 3);
    const int y = y0 + yl;
#pragma unroll
    for (int k = 0; k < QPL; ++k) {
        const int dz = 2 * k + (lane >> 4);
        const int z = a.g.z0 + bz * 8 + dz;
        const int zl = (a.out_layer0 + l) * 8 + dz;
        valid[k] = x < a.g.nx && y < a.g.ny && z < a.g.z1;
        qx[k] = wn_lattice_coord(a.g.ox, a.g.sx, x);
        qy[k] = wn_lattice_coord(a.g.oy, a.g.sy, y);
        qz[k] = wn_lattice_coord(a.g.oz, a.g.sz, z);
        oidx[k] = valid[k] ? ((int64_t)zl * a.part_ny + yl) * a.g.nx + x : -1;
    }
}

// Points: warp w of block b owns slots [(b*8 + w) * 32*QPL, +32*QPL); `stage` = 24*QPL float4 of shared memory per warp.
template <int QPL>
__device__ __forceinline__ void list_points(const QueryArgs& a, int64_t block, const int wid, float4* stage, float (&qx)[QPL],
                                            float (&qy)[QPL], float (&qz)[QPL], bool (&valid)[QPL], int64_t (&oidx)[QPL],
                                            const int tasks_per_block = kQueryWarps)
{
    const int lane = threadIdx.x & 31;
    const int64_t wbase = (block * tasks_per_block + wid) * (32 * QPL);
    const bool staged = a.perm == nullptr && a.q_aligned16 && wbase + 32 * QPL <= a.n;
    if (staged) {
        // 32*QPL points = 96*QPL floats = 24*QPL float4, contiguous and 16-byte aligned: vectorised, coalesced
        const float4* src = reinterpret_cast<const float4*>(a.q + 3 * wbase);
        __syncwarp(); // a previous use of this warp's stage (k_tile_query runs several tasks per warp) has been read
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
__global__ void __launch_bounds__(kQueryThreads, WN_Q_MIN_CTAS) k_query(const QueryArgs a)
{
    __shared__ float4 stage[GRID ? 1 : kQueryWarps * 24 * QPL];
    // (one block of 8 warp tasks per CTA: runs of blocks with a CTA-local task counter, the scheme that gives k_tile_query +5 %,
    // were measured 3 % SLOWER here -- the loop costs registers under the 48-register cap)
    float qx[QPL], qy[QPL], qz[QPL], acc[QPL];
    bool valid[QPL];
    int64_t oidx[QPL];
    if (GRID)
        grid_points<QPL>(a, (int)blockIdx.x, threadIdx.x >> 5, qx, qy, qz, valid, oidx);
    else
        list_points<QPL>(a, blockIdx.x, threadIdx.x >> 5, stage + (threadIdx.x >> 5) * 24 * QPL, qx, qy, qz, valid, oidx);
#pragma unroll
    for (int k = 0; k < QPL; ++k) acc[k] = 0.0f;
    TravCounters cnt;
    warp_traverse<QPL, STATS, false>(a.tree, a.beta2, qx, qy, qz, valid, acc, nullptr, 0, cnt);
    write_results<QPL>(a, oidx, acc);
    if (STATS) flush_counters(a, cnt);
}

// ---- tiled path: plan ------------------------------------------------------------------------------------------
// frontier words: entry | (has_mixed_ancestor << 30)
__device__ __forceinline__ void plan_push_kids(const int4 k4, int flag, int* front, int* count, int* overflow, bool& ovf)
{
    const int kid[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (kid[s] >= 0) {
            const int pos = atomicAdd(count, 1);
            if (pos < kTileFrontCap)
                front[pos] = kid[s] | (flag << 30);
            else
                *overflow = 1, ovf = true;
        }
    }
}

__device__ __forceinline__ void plan_append(int* list, int* count, int cap, int value, int* overflow, bool& ovf)
{
    const int pos = atomicAdd(count, 1);
    if (pos < cap)
        list[pos] = value;
    else
        *overflow = 1, ovf = true;
}

// ascending sort of s[0..n), n <= 64, by ONE warp: two keys per lane (elements lane and lane + 32), bitonic network with
// shuffles, no shared-memory traffic and no block barrier
__device__ __forceinline__ void plan_warp_sort64(int* s, int n)
{
    const int lane = threadIdx.x & 31;
    int a = lane < n ? s[lane] : 0x7fffffff;
    int b = lane + 32 < n ? s[lane + 32] : 0x7fffffff;
#pragma unroll
    for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j == 32) { // k == 64: partner is the other key of the same lane, direction ascending
                const int lo = min(a, b), hi = max(a, b);
                a = lo;
                b = hi;
            } else {
                const int pa = __shfl_xor_sync(kFull, a, j), pb = __shfl_xor_sync(kFull, b, j);
                const bool lower = (lane & j) == 0;
                const bool asc_a = (lane & k) == 0;        // element index = lane
                const bool asc_b = ((lane + 32) & k) == 0; // element index = lane + 32
                a = (lower == asc_a) ? min(a, pa) : max(a, pa);
                b = (lower == asc_b) ? min(b, pb) : max(b, pb);
            }
        }
    }
    if (lane < n) s[lane] = a;
    if (lane + 32 < n) s[lane + 32] = b;
}

// ascending sort of s[0..n), 64 < n <= 2 * kTileFrontCap, distinct keys, by the whole CTA: 64-key chunks by the warps'
// shuffle network, then log2(n/64) merge rounds in which every key finds its rank in the sibling run by binary search
// (ping-pong with tmp). ~4x fewer instructions and ~6x fewer barriers than a bitonic network at the usual n ~ 100-300.
__device__ __forceinline__ void plan_merge_sort(int* s, int* tmp, int n)
{
    const int wid = threadIdx.x >> 5;
    for (int c = wid * 64; c < n; c += kPlanWarps * 64) plan_warp_sort64(s + c, min(64, n - c));
    __syncthreads();
    int* src = s;
    int* dst = tmp;
    for (int L = 64; L < n; L <<= 1) {
        for (int i = threadIdx.x; i < n; i += kPlanThreads) {
            const int start = i & ~(L - 1);
            const int sib = start ^ L;
            const int sib_len = max(0, min(L, n - sib));
            const int key = src[i];
            int lo = 0, hi = sib_len;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (src[sib + mid] < key)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            dst[min(start, sib) + (i - start) + lo] = key;
        }
        __syncthreads();
        int* x = src;
        src = dst;
        dst = x;
    }
    if (src != s) {
        for (int i = threadIdx.x; i < n; i += kPlanThreads) s[i] = src[i];
        __syncthreads();
    }
}

// ---- hierarchical planning: blocks above the tiles -----------------------------------------------------------------------
// Neighbouring tiles redo the same top of the breadth-first classification and sample largely the same far records. A block of
// 2^k x 2^k x (2^k | 1) tiles does that part once: it classifies against ITS bounding sphere, starting from its parent block's
// entry list (the top level: from the root),
//   far for the whole block and smooth across it -> evaluated at the block's own 4^3 Chebyshev points, added to the parent's
//                                                   interpolant evaluated there (exact: a tensor cubic is reproduced by 4^3 points)
//   near for the whole block, internal            -> expanded
//   anything else                                 -> entry list of the block: its children classify it against their own sphere
// so a tile starts a few levels above its own scale with most of its far field already in the samples it inherits. Which
// records a point accepts is untouched (the per-point test stays the reference's); the classification of a record against a
// tile is the same as from the root, because every ancestor of a listed record is near for the whole block, hence for the tile.
constexpr int kBlockPassCap = 1024;

__device__ __forceinline__ float plan_parent_interp(const float* __restrict__ up, float px, float py, float pz)
{
    // up[0..63] samples, up[64..66] centre, up[67..69] 1 / half-extent of the parent's box
    float wx[4], wy[4], wz[4];
    cheb_weights((px - __ldg(up + 64)) * __ldg(up + 67), wx);
    cheb_weights((py - __ldg(up + 65)) * __ldg(up + 68), wy);
    cheb_weights((pz - __ldg(up + 66)) * __ldg(up + 69), wz);
    float acc = 0.0f;
#pragma unroll
    for (int kz = 0; kz < 4; ++kz) {
        float sz = 0.0f;
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            const float4 row = __ldg(reinterpret_cast<const float4*>(up) + kz * 4 + ky);
            sz += wy[ky] * (wx[0] * row.x + wx[1] * row.y + wx[2] * row.z + wx[3] * row.w);
        }
        acc += wz[kz] * sz;
    }
    return acc;
}

// CTA timeline for tools/cta_timeline.py: one branch on a kernel parameter per CTA when tracing is off.
__device__ __forceinline__ unsigned long long trace_clock()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_begin(const QueryArgs& a)
{
    if (a.trace && threadIdx.x == 0) {
        unsigned int sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        a.trace[(size_t)blockIdx.x * 4 + 0] = trace_clock();
        a.trace[(size_t)blockIdx.x * 4 + 2] = sm;
    }
}
__device__ __forceinline__ void trace_end(const QueryArgs& a)
{
    if (a.trace && threadIdx.x == 0) a.trace[(size_t)blockIdx.x * 4 + 1] = trace_clock();
}

__global__ void __launch_bounds__(kPlanThreads) k_plan_block(const QueryArgs a)
{
    __shared__ int s_front[2][kTileFrontCap];
    __shared__ int s_pass[kBlockPassCap];
    __shared__ int s_far[kTileFarCap];
    __shared__ float s_samp[kPlanWarps][kTileSamples];
    __shared__ int s_fcnt[kPlanMaxRounds + 1];
    __shared__ int s_cnt[4]; // 0 pass, 1 far, 2 overflow / bad, 3 parent usable
    __shared__ float s_geo[8];
    __shared__ long long s_off;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const WnTreeView& t = a.tree;
    const int X = (int)blockIdx.x % a.lvl_bx, Y = ((int)blockIdx.x / a.lvl_bx) % a.lvl_by, Z = (int)blockIdx.x / (a.lvl_bx * a.lvl_by);
    const int pidx = a.up_hdr ? ((Z >> a.up_zs) * a.up_by + (Y >> 1)) * a.up_bx + (X >> 1) : 0;
    bool my_ovf = false;
    trace_begin(a);
    if (tid < 4) s_cnt[tid] = 0;
    for (int r = tid; r <= kPlanMaxRounds; r += kPlanThreads) s_fcnt[r] = 0;
    if (tid == 0) s_off = 0;
    __syncthreads();
    if (tid == 0) {
        const int span = 8 << a.lvl_shift;
        int zt, y0;
        layer_geom(a, Z << a.lvl_zshift, zt, y0);
        const int zp = a.g.z0 + zt * 8;
        const float lx = wn_lattice_coord(a.g.ox, a.g.sx, X * span), hx = wn_lattice_coord(a.g.ox, a.g.sx, X * span + span - 1);
        const float ly = wn_lattice_coord(a.g.oy, a.g.sy, y0 + Y * span), hy = wn_lattice_coord(a.g.oy, a.g.sy, y0 + Y * span + span - 1);
        const float lz = wn_lattice_coord(a.g.oz, a.g.sz, zp), hz = wn_lattice_coord(a.g.oz, a.g.sz, zp + (8 << a.lvl_zshift) - 1);
        const float lo[3] = {fminf(lx, hx), fminf(ly, hy), fminf(lz, hz)}, hi[3] = {fmaxf(lx, hx), fmaxf(ly, hy), fmaxf(lz, hz)};
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
        s_geo[6] = sqrtf(r2) * 1.0001f + 1e-30f;
        s_geo[7] = 0.0f;
        if (!finite) s_cnt[2] = 1;
        bool from_parent = false;
        if (a.up_hdr) {
            const PlanBlockHeader ph = a.up_hdr[pidx];
            from_parent = !(ph.flags & kTileFallback) && ph.count <= kTileFrontCap;
            if (from_parent) {
                s_fcnt[0] = ph.count;
                s_off = ph.offset; // read below by everybody
            }
        }
        s_cnt[3] = from_parent ? 1 : 0;
        if (!from_parent && t.n_entries > 1) plan_push_kids(__ldg(t.kids), 0, s_front[0], &s_fcnt[0], &s_cnt[2], my_ovf);
        if (t.n_entries <= 1) s_cnt[2] = 1; // a single leaf: nothing to share, the tiles plan by themselves
    }
    __syncthreads();
    const bool from_parent = s_cnt[3] != 0;
    if (from_parent) {
        const int* src = reinterpret_cast<const int*>(a.plan_arena + s_off);
        const int n0 = s_fcnt[0];
        for (int j = tid; j < n0; j += kPlanThreads) s_front[0][j] = __ldg(src + j);
    }
    __syncthreads();
    const float cx = s_geo[0], cy = s_geo[1], cz = s_geo[2], ra = s_geo[6];
    const float hx = s_geo[3] > 0.0f ? 1.0f / s_geo[3] : 0.0f, hy = s_geo[4] > 0.0f ? 1.0f / s_geo[4] : 0.0f,
                hz = s_geo[5] > 0.0f ? 1.0f / s_geo[5] : 0.0f;
    // half extents of the box, inflated like the radius (relative 1e-4 of the radius on every axis)
    const float bhx = hx + 1e-4f * ra, bhy = hy + 1e-4f * ra, bhz = hz + 1e-4f * ra;
    int round = 0;
    bool stop = s_cnt[2] != 0;
    while (!stop) {
        const int F = min(s_fcnt[round], kTileFrontCap);
        if (F == 0) break;
        const int* cur = s_front[round & 1];
        int* nxt = s_front[(round + 1) & 1];
        int* ncnt = &s_fcnt[round + 1];
        for (int idx = tid; idx < F; idx += kPlanThreads) {
            const int e = cur[idx] & 0x3fffffff;
            const float4 f0 = rec_hot(t, e, 0);
            const bool leaf = __float_as_int(f0.w) < 0;
            const float thr = fabsf(f0.w) * a.beta2;
            const float dx = cx - f0.x, dy = cy - f0.y, dz = cz - f0.z;
            const float D = sqrtf(dx * dx + dy * dy + dz * dz);
            // nearest / farthest point of the block's BOX from the record's centre (the points lie inside the box; the bounding
            // sphere of a cube has 2.7x its volume and called many records "mixed" that no point of the box disagrees on);
            // 1e-4 relative slack on the threshold + an absolute one on the box for the rounding of the per-point test
            const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
            const float nx = fmaxf(ax - bhx, 0.0f), ny = fmaxf(ay - bhy, 0.0f), nz = fmaxf(az - bhz, 0.0f);
            const float fx = ax + bhx, fy = ay + bhy, fz = az + bhz;
            const bool allfar = nx * nx + ny * ny + nz * nz > thr * 1.0001f;
            const bool allnear = fx * fx + fy * fy + fz * fz <= thr * 0.9999f;
            if (allfar && D >= a.kappa * ra && D - sqrtf(fabsf(f0.w)) >= 0.5f * a.kappa * ra)
                plan_append(s_far, &s_cnt[1], kTileFarCap, e, &s_cnt[2], my_ovf);
            else if (allnear && !leaf)
                plan_push_kids(__ldg(t.kids + e), 0, nxt, ncnt, &s_cnt[2], my_ovf);
            else
                plan_append(s_pass, &s_cnt[0], kBlockPassCap, e, &s_cnt[2], my_ovf);
        }
        ++round;
        if (round >= kPlanMaxRounds) {
            my_ovf = true;
            if (tid == 0) s_cnt[2] = 1;
        }
        stop = __syncthreads_or((int)my_ovf) != 0;
    }
    __syncthreads();
    bool fallback = s_cnt[2] != 0;
    const int n_pass = fallback ? 0 : s_cnt[0], n_far = fallback ? 0 : s_cnt[1];
    if (tid == 0 && !fallback && n_pass > 0) {
        const long long bytes = ((long long)n_pass * 4 + 15) & ~15ll;
        const long long off = (long long)atomicAdd(a.plan_cursor, (unsigned long long)bytes);
        if (off + bytes > a.plan_arena_bytes) s_cnt[2] = 1;
        s_off = off;
    }
    __syncthreads();
    fallback = s_cnt[2] != 0;
    if (!fallback) {
        int* dst = reinterpret_cast<int*>(a.plan_arena + s_off);
        for (int j = tid; j < n_pass; j += kPlanThreads) dst[j] = s_pass[j];
    }
    // far set at the block's Chebyshev points, on top of what the parent already holds
    float sacc[2] = {0.0f, 0.0f};
    bool bad = false;
    if (!fallback) {
        float px[2], py[2], pz[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int sidx = lane + 32 * k;
            px[k] = cx + hx * cheb_node(sidx & 3);
            py[k] = cy + hy * cheb_node((sidx >> 2) & 3);
            pz[k] = cz + hz * cheb_node(sidx >> 4);
        }
        if (wid == 0 && from_parent) {
            const float* up = a.up_samples + (int64_t)pidx * kTileSampleStride;
#pragma unroll
            for (int k = 0; k < 2; ++k) sacc[k] = plan_parent_interp(up, px[k], py[k], pz[k]);
        }
        for (int m = wid; m < n_far; m += kPlanWarps) {
            const int e = s_far[m];
            const float4 f0 = rec_hot(t, e, 0), f1 = rec_hot(t, e, 1), f2 = rec_cold(t, e, 0), f3 = rec_cold(t, e, 1),
                         f4 = rec_cold(t, e, 2), f5 = rec_cold(t, e, 3);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float rx = px[k] - f0.x, ry = py[k] - f0.y, rz = pz[k] - f0.z;
                const float l2 = rx * rx + ry * ry + rz * rz;
                const float om = eval_record(rx, ry, rz, l2, f1, f2, f3, f4, f5);
                bad = bad || !(fabsf(om) <= 3.402823466e38f);
                sacc[k] += om;
            }
        }
        bad = bad || !(fabsf(sacc[0]) <= 3.402823466e38f) || !(fabsf(sacc[1]) <= 3.402823466e38f);
        s_samp[wid][lane] = sacc[0];
        s_samp[wid][lane + 32] = sacc[1];
    }
    fallback = __syncthreads_or((int)(bad || fallback)) != 0;
    float* sout = a.lvl_samples + (int64_t)blockIdx.x * kTileSampleStride;
    for (int j = tid; !fallback && j < kTileSamples; j += kPlanThreads) {
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < kPlanWarps; ++w) v += s_samp[w][j];
        sout[j] = v;
    }
    if (tid < 8) sout[kTileSamples + tid] = s_geo[tid];
    if (tid == 0) {
        PlanBlockHeader h;
        h.offset = s_off;
        h.count = fallback ? 0 : n_pass;
        h.flags = fallback ? kTileFallback : 0;
        a.lvl_hdr[blockIdx.x] = h;
        if (a.stats) atomicAdd(a.stats + 1, (unsigned long long)(fallback ? 0 : n_far) * kTileSamples);
        trace_end(a);
    }
}

template <bool GRID>
__global__ void WN_PLAN_BOUNDS k_tile_plan(const QueryArgs a)
{
    __shared__ int s_front[2][kTileFrontCap];
    __shared__ int s_cond[kTileAllCap];
    __shared__ int s_dir[kTileDirCap];
    __shared__ int s_exact[kTileExactCap];
    __shared__ int s_far[kTileFarCap];
    __shared__ float s_samp[kPlanWarps][kTileSamples];
    __shared__ int s_fcnt[kPlanMaxRounds + 1]; // frontier size of every round
    __shared__ int s_cnt[8]; // 0,1 unused; 2 cond; 3 far; 4 overflow / bad; 5 dir; 6 exact; 7 triangles
    __shared__ float s_geo[8];
    __shared__ float s_red[kPlanWarps][6];
    __shared__ long long s_off;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const WnTreeView& t = a.tree;
    // probe mode (probe_stride > 0): classify every probe_stride-th tile only and add up the class sizes, so that the host
    // can decide whether the tiled path pays off for this batch (it does when the far set is a large share of the work)
    const int tile = a.probe_stride > 0 ? (int)blockIdx.x * a.probe_stride : (int)blockIdx.x;
    trace_begin(a);
    int pidx = 0; // parent block of this tile (hierarchical planning, lattices only)
    if (GRID && a.up_hdr) {
        const int tbx = tile % a.tiles_x, tby = (tile / a.tiles_x) % a.tiles_y, tl = tile / (a.tiles_x * a.tiles_y);
        pidx = ((tl >> a.up_zs) * a.up_by + (tby >> 1)) * a.up_bx + (tbx >> 1);
    }

    // ---- tile bounding sphere ------------------------------------------------------------------------------------
    if (GRID) {
        if (tid == 0) {
            const int bx = tile % a.tiles_x;
            const int by = (tile / a.tiles_x) % a.tiles_y;
            int bz, y0;
            layer_geom(a, tile / (a.tiles_x * a.tiles_y), bz, y0);
            const float lx = wn_lattice_coord(a.g.ox, a.g.sx, bx * 8), hx = wn_lattice_coord(a.g.ox, a.g.sx, bx * 8 + 7);
            const float ly = wn_lattice_coord(a.g.oy, a.g.sy, y0 + by * 8), hy = wn_lattice_coord(a.g.oy, a.g.sy, y0 + by * 8 + 7);
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
        const int64_t base = ((int64_t)tile + a.tile_base) * kTileQueries;
        for (int k = tid; k < kTileQueries; k += kPlanThreads) {
            const int64_t s = base + k;
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
    bool my_ovf = false; // this thread overflowed a list: decided at the barrier's OR, never by reading the shared flag mid-round
    if (tid < 8) s_cnt[tid] = 0;
    for (int r = tid; r <= kPlanMaxRounds; r += kPlanThreads) s_fcnt[r] = 0;
    if (tid == 0) s_off = 0;
    __syncthreads();
    if (tid == 0) {
        float lo[3], hi[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            lo[d] = s_red[0][d];
            hi[d] = s_red[0][3 + d];
            if (!GRID) {
                for (int w = 1; w < kPlanWarps; ++w) {
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
        // hierarchical planning: start from the parent block's entry list (k_plan_block) instead of the root
        bool from_parent = false;
        if (GRID && a.up_hdr && a.probe_stride == 0 && n_entries > 1) {
            const PlanBlockHeader ph = a.up_hdr[pidx];
            from_parent = !(ph.flags & kTileFallback) && ph.count <= kTileFrontCap;
            if (from_parent) {
                s_fcnt[0] = ph.count;
                s_off = ph.offset;
            }
        }
        s_cnt[0] = from_parent ? 1 : 0;
        if (from_parent) {
        } else if (n_entries > 1) {
            plan_push_kids(__ldg(t.kids), 0, s_front[0], &s_fcnt[0], &s_cnt[4], my_ovf);
        } else if (n_entries == 1) {
            s_exact[0] = 0; // the root is a leaf: exact for everybody
            s_cnt[6] = 1;
        }
    }
    __syncthreads();
    const bool from_parent = GRID && s_cnt[0] != 0;
    if (from_parent) {
        const int* src = reinterpret_cast<const int*>(a.plan_arena + s_off);
        const int n0 = s_fcnt[0];
        for (int j = tid; j < n0; j += kPlanThreads) s_front[0][j] = __ldg(src + j);
        __syncthreads();
        if (tid == 0) s_off = 0;
    }
    const float cx = s_geo[0], cy = s_geo[1], cz = s_geo[2], ra = s_geo[6];
    const float hx = s_geo[3] > 0.0f ? 1.0f / s_geo[3] : 0.0f, hy = s_geo[4] > 0.0f ? 1.0f / s_geo[4] : 0.0f,
                hz = s_geo[5] > 0.0f ? 1.0f / s_geo[5] : 0.0f;
    // half extents of the box, inflated like the radius (relative 1e-4 of the radius on every axis)
    const float bhx = hx + 1e-4f * ra, bhy = hy + 1e-4f * ra, bhz = hz + 1e-4f * ra;

    // ---- breadth-first classification ----------------------------------------------------------------------------
    // One barrier per round: every round has its own frontier counter (no reset between rounds), and the decision to stop
    // is taken by the barrier itself (an OR over all threads), so no thread can leave the loop on a flag another thread
    // is still about to set.
    int round = 0;
    bool stop = s_cnt[4] != 0;
    while (!stop) {
        const int F = min(s_fcnt[round], kTileFrontCap);
        if (F == 0) break;
        const int* cur = s_front[round & 1];
        int* nxt = s_front[(round + 1) & 1];
        int* ncnt = &s_fcnt[round + 1];
        for (int idx = tid; idx < F; idx += kPlanThreads) {
            const int word = cur[idx];
            const int e = word & 0x3fffffff, manc = (word >> 30) & 1;
            const float4 f0 = rec_hot(t, e, 0);
            const int4 k4 = __ldg(t.kids + e);
            const bool leaf = __float_as_int(f0.w) < 0;
            const float thr = fabsf(f0.w) * a.beta2;
            const float dx = cx - f0.x, dy = cy - f0.y, dz = cz - f0.z;
            const float D = sqrtf(dx * dx + dy * dy + dz * dz);
            // nearest / farthest point of the block's BOX from the record's centre (the points lie inside the box; the bounding
            // sphere of a cube has 2.7x its volume and called many records "mixed" that no point of the box disagrees on);
            // 1e-4 relative slack on the threshold + an absolute one on the box for the rounding of the per-point test
            const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
            const float nx = fmaxf(ax - bhx, 0.0f), ny = fmaxf(ay - bhy, 0.0f), nz = fmaxf(az - bhz, 0.0f);
            const float fx = ax + bhx, fy = ay + bhy, fz = az + bhz;
            const bool allfar = nx * nx + ny * ny + nz * nz > thr * 1.0001f;
            const bool allnear = fx * fx + fy * fy + fz * fz <= thr * 0.9999f;
            if (allfar) {
                // far set: the record's field must be smooth across the tile, i.e. the tile is small against its distance
                // both to the expansion centre and to the nearest possible source point (bounding sphere of radius R)
                if (manc)
                    plan_append(s_cond, &s_cnt[2], kTileAllCap, tile_key(e, leaf, kClsCondFar), &s_cnt[4], my_ovf);
                else if (D >= a.kappa * ra && D - sqrtf(fabsf(f0.w)) >= 0.5f * a.kappa * ra)
                    plan_append(s_far, &s_cnt[3], kTileFarCap, e, &s_cnt[4], my_ovf);
                else
                    plan_append(s_dir, &s_cnt[5], kTileDirCap, e, &s_cnt[4], my_ovf);
            } else if (allnear) {
                if (!leaf)
                    plan_push_kids(k4, manc, nxt, ncnt, &s_cnt[4], my_ovf);
                else if (manc)
                    plan_append(s_cond, &s_cnt[2], kTileAllCap, tile_key(e, true, kClsCond), &s_cnt[4], my_ovf);
                else
                    plan_append(s_exact, &s_cnt[6], kTileExactCap, e, &s_cnt[4], my_ovf);
            } else {
                plan_append(s_cond, &s_cnt[2], kTileAllCap, tile_key(e, leaf, kClsCond), &s_cnt[4], my_ovf);
                if (!leaf) plan_push_kids(k4, 1, nxt, ncnt, &s_cnt[4], my_ovf);
            }
        }
        ++round;
        if (round >= kPlanMaxRounds) { // deeper than any sane hierarchy: generic path
            my_ovf = true;
            if (tid == 0) s_cnt[4] = 1;
        }
        stop = __syncthreads_or((int)my_ovf) != 0;
    }
    __syncthreads();
    bool fallback = s_cnt[4] != 0;
    const int n_cond = fallback ? 0 : s_cnt[2], n_far = fallback ? 0 : s_cnt[3], n_dir = fallback ? 0 : s_cnt[5], n_ex = fallback ? 0 : s_cnt[6];
    if (a.probe_stride > 0) {
        if (tid == 0) {
            atomicAdd(a.probe + 0, (unsigned long long)n_far);
            atomicAdd(a.probe + 1, (unsigned long long)n_cond);
            atomicAdd(a.probe + 2, (unsigned long long)n_dir);
            atomicAdd(a.probe + 3, (unsigned long long)n_ex);
            atomicAdd(a.probe + 4, (unsigned long long)(fallback ? 1 : 0));
        }
        return;
    }

    // ---- back to depth-first order. The atomics above appended in arbitrary order; sorting restores the order the skip
    //      links need and makes every sum deterministic. Short lists (the usual case) are sorted by one warp each, in
    //      parallel, with shuffles; long ones by the whole CTA. -------------------------------------------------------
    int* const lists[4] = {s_cond, s_far, s_dir, s_exact};
    const int lens[4] = {n_cond, n_far, n_dir, n_ex};
    for (int l = wid; l < 4; l += kPlanWarps)
        if (lens[l] > 1 && lens[l] <= 64) plan_warp_sort64(lists[l], lens[l]);
    __syncthreads();
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        if (lens[l] > 64) plan_merge_sort(lists[l], &s_front[0][0], lens[l]); // uniform over the CTA; the frontier is done
    }

    // ---- triangles of the exact class: exclusive prefix of the leaf sizes (warp 0) --------------------------------------
    int* s_exoff = s_front[0]; // the frontier is done; reuse it for the triangle offsets of the exact leaves
    if (wid == 0) {
        int run = 0;
        for (int base = 0; base < n_ex; base += 32) {
            const int j = base + lane;
            const int c = j < n_ex ? (rec_link(t, s_exact[j]) & (WN_MAX_LEAF_SIZE - 1)) + 1 : 0;
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(kFull, inc, o);
                if (lane >= o) inc += v;
            }
            if (j < n_ex) s_exoff[j] = run + inc - c;
            run += __shfl_sync(kFull, inc, 31);
        }
        if (lane == 0) s_cnt[7] = run;
    }
    __syncthreads();
    const int n_tri = s_cnt[7];

    // ---- packet allocation and contents ---------------------------------------------------------------------------------
    const long long cond_bytes = (long long)n_cond * 32;
    const long long bytes = cond_bytes + (long long)n_dir * 96 + (long long)n_tri * 48;
    if (tid == 0 && !fallback && bytes > 0) {
        const long long off = (long long)atomicAdd(a.plan_cursor, (unsigned long long)bytes);
        if (off + bytes > a.plan_arena_bytes) s_cnt[4] = 1; // arena exhausted: generic path for this tile
        s_off = off;
    }
    __syncthreads();
    fallback = s_cnt[4] != 0;
    if (!fallback) {
        char* pk = a.plan_arena + s_off;
        float4* out_cond = reinterpret_cast<float4*>(pk);
        float4* out_dir = reinterpret_cast<float4*>(pk + cond_bytes);
        float4* out_tri = out_dir + (long long)n_dir * 6;
        for (int j = tid; j < n_cond; j += kPlanThreads) {
            // skip link in list coordinates: first conditional record at or after the end of this record's subtree
            const int key = s_cond[j], e = key >> 3;
            const float4 f0 = rec_hot(t, e, 0);
            const int lk = rec_link(t, e);
            const int end = (key & 4) ? e + 1 : lk;
            const int target = end << 3;
            int lo = j + 1, hi = n_cond;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_cond[mid] < target)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            // self-contained item: centre + the acceptance threshold beta^2 R^2 exactly as the traversal forms it, key, skip
            // position, leaf link
            out_cond[2 * j] = make_float4(f0.x, f0.y, f0.z, __fmul_rn(fabsf(f0.w), a.beta2));
            out_cond[2 * j + 1] = make_float4(__int_as_float(key), __int_as_float(lo), __int_as_float(lk), 0.0f);
        }
        for (int j = tid; j < n_dir * 6; j += kPlanThreads) {
            const int e = s_dir[j / 6], r = j % 6;
            out_dir[j] = r < 2 ? rec_hot(t, e, r) : rec_cold(t, e, r - 2);
        }
        for (int j = tid; j < n_ex; j += kPlanThreads) {
            const int lk = rec_link(t, s_exact[j]);
            const int first = lk >> WN_LEAF_COUNT_BITS, count = (lk & (WN_MAX_LEAF_SIZE - 1)) + 1;
            float4* dst = out_tri + (long long)s_exoff[j] * 3;
            for (int r = 0; r < 3 * count; ++r) dst[r] = __ldg(t.tri + 3 * (int64_t)first + r);
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
        if (GRID && from_parent && wid == 0) {
            // what the blocks above already sampled (their far sets), interpolated to this tile's Chebyshev points
            const float* up = a.up_samples + (int64_t)pidx * kTileSampleStride;
#pragma unroll
            for (int k = 0; k < 2; ++k) sacc[k] = plan_parent_interp(up, px[k], py[k], pz[k]);
        }
        for (int m = wid; m < n_far; m += kPlanWarps) {
            const int e = s_far[m];
            const float4 f0 = rec_hot(t, e, 0), f1 = rec_hot(t, e, 1), f2 = rec_cold(t, e, 0), f3 = rec_cold(t, e, 1),
                         f4 = rec_cold(t, e, 2), f5 = rec_cold(t, e, 3);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float rx = px[k] - f0.x, ry = py[k] - f0.y, rz = pz[k] - f0.z;
                const float l2 = rx * rx + ry * ry + rz * rz;
                const float om = eval_record(rx, ry, rz, l2, f1, f2, f3, f4, f5);
                bad = bad || !(fabsf(om) <= 3.402823466e38f);
                sacc[k] += om;
            }
        }
        s_samp[wid][lane] = sacc[0];
        s_samp[wid][lane + 32] = sacc[1];
    }
    fallback = __syncthreads_or((int)(bad || fallback)) != 0;
    float* sout = a.plan_samples + (int64_t)blockIdx.x * kTileSampleStride;
    for (int j = tid; !fallback && j < kTileSamples; j += kPlanThreads) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < kPlanWarps; ++w) s += s_samp[w][j];
        sout[j] = s;
    }
    if (tid < 8) sout[kTileSamples + tid] = s_geo[tid];
    if (tid == 0) {
        TileHeader h;
        h.n_cond = fallback ? 0 : n_cond;
        h.n_dir = fallback ? 0 : n_dir;
        h.n_tri = fallback ? 0 : n_tri;
        h.flags = fallback ? kTileFallback : 0;
        h.offset = s_off;
        h.pad = 0;
        a.plan_hdr[blockIdx.x] = h;
        // Longest-processing-time-first: the query kernel takes heavy tiles (long conditional walk, or generic fallback) before
        // light ones, so the tail of the launch is made of short CTAs. Heavy tiles fill the order from the front, light ones
        // from the back; within a class the order stays close to the launch order (neighbouring tiles share records in L1/L2).
        unsigned int* cnt = reinterpret_cast<unsigned int*>(a.plan_cursor + 1);
        const bool heavy = fallback || n_cond >= a.heavy_cond;
        const int pos = heavy ? (int)atomicAdd(cnt, 1u) : a.launch_tiles - 1 - (int)atomicAdd(cnt + 1, 1u);
        a.tile_order[pos] = (int)blockIdx.x;
        if (a.stats) {
            // executed work of the plan: far-set evaluations at the sample points (counted as far-field evaluations)
            atomicAdd(a.stats + 1, (unsigned long long)(fallback ? 0 : n_far) * kTileSamples);
        }
        trace_end(a);
    }
}

// ---- tiled path: query -----------------------------------------------------------------------------------------
// A CTA owns a run of consecutive tiles; its warps pull (tile, 4x4x4 sub-block) tasks from a CTA-local counter, so a warp
// whose sub-block lies close to the surface (long conditional walk) does not hold seven finished warps at a barrier:
// nothing here synchronises the CTA after the counter is set up. The plan's packets are read through L1 (warp-uniform).
template <bool GRID, bool STATS>
__global__ void __launch_bounds__(kQueryThreads, WN_TQ_MIN_CTAS) k_tile_query(const QueryArgs a)
{
    __shared__ int s_next;
    __shared__ float4 stage[GRID ? 1 : kQueryWarps * 24 * kTileQPL];
    constexpr int QPL = kTileQPL;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_next = 0;
    trace_begin(a);
    __syncthreads();
    // runs of tiles_per_cta tiles, except at the end of the launch (the light end of the heavy-first order), where a CTA takes one
    // tile: the launch tail is as long as its last CTAs, and those are then the shortest ones
    const bool in_run = (int)blockIdx.x < a.run_ctas;
    const int tile0 = in_run ? (int)blockIdx.x * a.tiles_per_cta : a.run_ctas * a.tiles_per_cta + ((int)blockIdx.x - a.run_ctas);
    const int n_task = min(in_run ? a.tiles_per_cta : 1, a.launch_tiles - tile0) * kTileTasks;
    TravCounters cnt;
    while (true) {
        int task = 0;
        if (lane == 0) task = atomicAdd(&s_next, 1);
        task = __shfl_sync(kFull, task, 0);
        if (task >= n_task) break;
        const int tile = __ldg(a.tile_order + tile0 + task / kTileTasks), sub = task % kTileTasks;
        float qx[QPL], qy[QPL], qz[QPL], acc[QPL];
        bool valid[QPL];
        int64_t oidx[QPL];
        if (GRID && QPL == 2)
            grid_points<QPL>(a, tile, sub, qx, qy, qz, valid, oidx);
        else if (GRID)
            tile_column_points<QPL>(a, tile, sub, qx, qy, qz, valid, oidx);
        else
            list_points<QPL>(a, (int64_t)tile + a.tile_base, sub, stage + (threadIdx.x >> 5) * 24 * QPL, qx, qy, qz, valid, oidx, kTileTasks);
        const int4 h4 = __ldg(reinterpret_cast<const int4*>(a.plan_hdr + tile));
        const long long offset = __ldg(&a.plan_hdr[tile].offset);
        const int n_cond = h4.x, n_dir = h4.y, n_tri = h4.z;
        const bool fallback = (h4.w & kTileFallback) != 0;
        const char* pk = a.plan_arena + offset;
        bool bad = false;
#pragma unroll
        for (int k = 0; k < QPL; ++k) acc[k] = 0.0f;
        if (!fallback) {
            // ---- direct records: far for every point of the tile; gathered contiguously by the plan -----------------
            const float4* __restrict__ dr = reinterpret_cast<const float4*>(pk + (long long)n_cond * 32);
            for (int j = 0; j < n_dir; ++j) {
                const float4 f0 = __ldg(dr + 0), f1 = __ldg(dr + 1), f2 = __ldg(dr + 2), f3 = __ldg(dr + 3), f4 = __ldg(dr + 4), f5 = __ldg(dr + 5);
                dr += 6;
#pragma unroll
                for (int k = 0; k < QPL; ++k) {
                    const float rx = qx[k] - f0.x, ry = qy[k] - f0.y, rz = qz[k] - f0.z;
                    const float l2 = __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz)); // as in warp_traverse
                    acc[k] += eval_record(rx, ry, rz, l2, f1, f2, f3, f4, f5); // non-finite values are caught after the walk
                }
            }
            // ---- exact triangles: leaves that are near for every point of the tile --------------------------------------
            const float4* __restrict__ tr = dr; // triangles follow the direct records in the packet
            for (int j = 0; j < n_tri; ++j) {
                const float4 ta = __ldg(tr + 0), tb = __ldg(tr + 1), tc = __ldg(tr + 2);
                tr += 3;
#pragma unroll
                for (int k = 0; k < QPL; ++k) acc[k] += wn_tri_solid_angle(qx[k], qy[k], qz[k], ta, tb, tc);
            }
            if (STATS) {
#pragma unroll
                for (int k = 0; k < QPL; ++k) {
                    if (valid[k]) {
                        cnt.A += n_dir;
                        cnt.E += n_tri;
                    }
                }
            }
            // ---- conditional records --------------------------------------------------------------------------------
            bad = tile_cond_walk<QPL, STATS>(a.tree, qx, qy, qz, valid, acc, reinterpret_cast<const float4*>(pk), n_cond, cnt) || bad;
            // the direct records are not checked one by one: a non-finite value poisons the sum for good (inf stays inf or turns NaN)
#pragma unroll
            for (int k = 0; k < QPL; ++k) bad = bad || (oidx[k] >= 0 && !(fabsf(acc[k]) <= 3.402823466e38f));
        }
        // The fallback is a per-warp matter: a point's result never depends on its neighbours.
        if (fallback || __any_sync(kFull, bad)) {
            // generic traversal of the whole tree for this sub-block (list overflow, non-finite far field, degenerate tile)
#pragma unroll
            for (int k = 0; k < QPL; ++k) acc[k] = 0.0f;
            warp_traverse<QPL, STATS, false>(a.tree, a.beta2, qx, qy, qz, valid, acc, nullptr, 0, cnt);
        } else {
            const float4* __restrict__ samp = reinterpret_cast<const float4*>(a.plan_samples + (int64_t)tile * kTileSampleStride);
            const float4 g0 = __ldg(samp + 16), g1 = __ldg(samp + 17); // centre xyz, 1/half-extent xyz
            float wx[QPL][4], wy[QPL][4], wz[QPL][4], far[QPL];
#pragma unroll
            for (int k = 0; k < QPL; ++k) {
                cheb_weights((qx[k] - g0.x) * g0.w, wx[k]);
                cheb_weights((qy[k] - g0.y) * g1.x, wy[k]);
                cheb_weights((qz[k] - g0.z) * g1.y, wz[k]);
                far[k] = 0.0f;
            }
#pragma unroll
            for (int kz = 0; kz < 4; ++kz) {
                float sz[QPL];
#pragma unroll
                for (int k = 0; k < QPL; ++k) sz[k] = 0.0f;
#pragma unroll
                for (int ky = 0; ky < 4; ++ky) {
                    const float4 row = __ldg(samp + kz * 4 + ky);
#pragma unroll
                    for (int k = 0; k < QPL; ++k)
                        sz[k] += wy[k][ky] * (wx[k][0] * row.x + wx[k][1] * row.y + wx[k][2] * row.z + wx[k][3] * row.w);
                }
#pragma unroll
                for (int k = 0; k < QPL; ++k) far[k] += wz[k][kz] * sz[k];
            }
#pragma unroll
            for (int k = 0; k < QPL; ++k) acc[k] += far[k];
        }
        write_results<QPL>(a, oidx, acc);
    }
    if (STATS) flush_counters(a, cnt);
    if (a.trace) {
        __syncthreads();
        trace_end(a);
    }
}

} // namespace wn
