// wn_query.cuh — sm_100a query kernels.
//
// K6 k_query : batched tree query. Replaces UT_SolidAngle::computeSolidAngle (modules/winding/src/FastWindingNumber.cpp:
//              66,75) for a whole batch. One warp owns 32*QPL spatially adjacent queries and walks the
//              depth-first record array once for all of them (stackless: "descend" = i+1, "skip subtree" = link[i]).
//              Every lane keeps the reference's per-point semantics through a private resume index: a lane that
//              accepted a far-field record ignores entries until the end of that subtree, lanes that must descend
//              keep going. The warp only leaves a subtree when no lane needs it. Far field = folded order-2 Taylor
//              record (23 floats), leaves = exact Van Oosterom-Strackee triangles. Node reads are warp-uniform
//              broadcasts of float4 SoA arrays; query points are generated from the lattice or loaded vectorised
//              (float4) through shared memory.
// K7 k_exact : exact all-pairs mode. Triangles staged in shared memory tiles, one query per thread, tile partial sums
//              added with compensated summation; k_exact_small splits the triangles of one query over a warp and
//              reduces with shuffles (small batches). Partial sums over triangle chunks are combined by k_exact_reduce.
// K9 k_point_bounds / k_point_morton : Morton keys of incoherent query sets for wn::radix_sort_pairs.
// All FP32 CUDA-core work (FMA pipe + MUFU rsqrt/atan): no tensor cores by design (BASELINE.json north_star).
#pragma once

#include <cuda_runtime.h>

#include "wn_device.cuh"
#include "wn_build.cuh"

namespace wn {

constexpr int kQueryThreads = 256;
constexpr unsigned kFull = 0xffffffffu;

struct GridDesc
{
    float ox, oy, oz, sx, sy, sz;
    int nx, ny, nz;
    int z0, z1; // slab
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
    // grid mode
    GridDesc g;
    int tiles_x, tiles_y;
    // outputs (either may be null)
    float* out_omega;
    uint8_t* out_inside;
    unsigned long long* stats; // [4] tests, approx, exact, lane slots
};

template <int QPL, bool GRID, bool STATS>
__global__ void __launch_bounds__(kQueryThreads) k_query(const QueryArgs a)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n_entries = a.tree.n_entries;
    float qx[QPL], qy[QPL], qz[QPL], acc[QPL];
    int skip[QPL];
    int64_t oidx[QPL];

    if (GRID) {
        const int bx = blockIdx.x % a.tiles_x;
        const int by = (blockIdx.x / a.tiles_x) % a.tiles_y;
        const int bz = blockIdx.x / (a.tiles_x * a.tiles_y);
        const int x = bx * 8 + (wid & 1) * 4 + (lane & 3);
        const int y = by * 8 + ((wid >> 1) & 1) * 4 + ((lane >> 2) & 3);
#pragma unroll
        for (int k = 0; k < QPL; ++k) {
            const int z = a.g.z0 + bz * (4 * QPL) + (wid >> 2) * (2 * QPL) + 2 * k + (lane >> 4);
            const bool valid = x < a.g.nx && y < a.g.ny && z < a.g.z1;
            qx[k] = wn_lattice_coord(a.g.ox, a.g.sx, x);
            qy[k] = wn_lattice_coord(a.g.oy, a.g.sy, y);
            qz[k] = wn_lattice_coord(a.g.oz, a.g.sz, z);
            oidx[k] = valid ? ((int64_t)(z - a.g.z0) * a.g.ny + y) * a.g.nx + x : -1;
            skip[k] = valid ? 0 : n_entries;
            acc[k] = 0.0f;
        }
    } else {
        __shared__ float4 stage[kQueryThreads / 32][24 * QPL];
        const int64_t wbase = ((int64_t)blockIdx.x * (kQueryThreads / 32) + wid) * (32 * QPL);
        const bool staged = a.perm == nullptr && a.q_aligned16 && wbase + 32 * QPL <= a.n;
        if (staged) {
            // 32*QPL points = 96*QPL floats = 24*QPL float4, contiguous and 16-byte aligned
            const float4* src = reinterpret_cast<const float4*>(a.q + 3 * wbase);
            for (int j = lane; j < 24 * QPL; j += 32) stage[wid][j] = __ldg(src + j);
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < QPL; ++k) {
            const int64_t s = wbase + k * 32 + lane;
            const bool valid = s < a.n;
            int64_t p = -1;
            qx[k] = qy[k] = qz[k] = 0.0f;
            if (valid) {
                p = a.perm ? (int64_t)a.perm[s] : s;
                if (staged) {
                    const float* f = reinterpret_cast<const float*>(&stage[wid][0]) + 3 * (k * 32 + lane);
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
            skip[k] = valid ? 0 : n_entries;
            acc[k] = 0.0f;
        }
    }

    unsigned long long cT = 0, cA = 0, cE = 0, cV = 0;
    const float4* __restrict__ r0 = a.tree.rec[0];
    const float4* __restrict__ r1 = a.tree.rec[1];
    const float4* __restrict__ r2 = a.tree.rec[2];
    const float4* __restrict__ r3 = a.tree.rec[3];
    const float4* __restrict__ r4 = a.tree.rec[4];
    const float4* __restrict__ r5 = a.tree.rec[5];
    const int* __restrict__ link = a.tree.link;
    const float4* __restrict__ tris = a.tree.tri;

    // entry 0 is the root, which is never approximated (A.5): start at its first child unless the root is itself a leaf
    int i = n_entries > 1 ? 1 : 0;
    while (i < n_entries) {
        const float4 f0 = __ldg(r0 + i);
        const int lk = __ldg(link + i);
        const bool leaf = __float_as_int(f0.w) < 0;
        // The accept/descend decision is formed exactly like the reference forms it (unfused |r|^2 <= R^2 * beta^2,
        // SURVEY.md A.5), so a point takes the same branch at every node as in the CPU algorithm: a decision flipped by
        // an FMA's single rounding would change Omega by that node's whole truncation error (~1e-4 * 4 pi at beta = 2).
        const float thr = __fmul_rn(fabsf(f0.w), a.beta2);
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
            const bool nr = l2[k] <= thr;
            nearq[k] = active && nr;
            farq[k] = active && !nr;
            anyfar |= farq[k];
            if (STATS) cT += active ? 1 : 0;
        }
        if (STATS) cV += (lane == 0) ? 32 * QPL : 0;
        if (__any_sync(kFull, anyfar)) {
            const float4 f1 = __ldg(r1 + i), f2 = __ldg(r2 + i), f3 = __ldg(r3 + i), f4 = __ldg(r4 + i), f5 = __ldg(r5 + i);
            const int after = leaf ? i + 1 : lk;
#pragma unroll
            for (int k = 0; k < QPL; ++k) {
                if (farq[k]) {
                    const float om = wn_eval_record(rx[k], ry[k], rz[k], l2[k], f1, f2, f3, f4, f5);
                    if (fabsf(om) <= 3.402823466e38f) {
                        acc[k] += om;
                        skip[k] = after;
                        if (STATS) ++cA;
                    } else {
                        nearq[k] = true; // non-finite expansion: descend instead (A.5)
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
                for (int t = 0; t < count; ++t) {
                    const float4 ta = __ldg(tris + 3 * (int64_t)(first + t));
                    const float4 tb = __ldg(tris + 3 * (int64_t)(first + t) + 1);
                    const float4 tc = __ldg(tris + 3 * (int64_t)(first + t) + 2);
#pragma unroll
                    for (int k = 0; k < QPL; ++k) {
                        if (nearq[k]) {
                            acc[k] += wn_tri_solid_angle(qx[k], qy[k], qz[k], ta, tb, tc);
                            if (STATS) ++cE;
                        }
                    }
                }
            }
            i = i + 1;
        } else {
            i = anynear ? i + 1 : lk;
        }
    }

#pragma unroll
    for (int k = 0; k < QPL; ++k) {
        if (oidx[k] >= 0) {
            if (a.out_omega) a.out_omega[oidx[k]] = acc[k];
            if (a.out_inside) a.out_inside[oidx[k]] = wn_inside_from_omega(acc[k]) ? 1 : 0;
        }
    }
    if (STATS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cT += __shfl_xor_sync(kFull, cT, o);
            cA += __shfl_xor_sync(kFull, cA, o);
            cE += __shfl_xor_sync(kFull, cE, o);
            cV += __shfl_xor_sync(kFull, cV, o);
        }
        if (lane == 0) {
            atomicAdd(a.stats + 0, cT);
            atomicAdd(a.stats + 1, cA);
            atomicAdd(a.stats + 2, cE);
            atomicAdd(a.stats + 3, cV);
        }
    }
}

// ---- K7 exact mode ---------------------------------------------------------------------------------------------
constexpr int kExactTile = 256; // triangles per shared-memory tile (12 KB)

struct ExactArgs
{
    const float4* tris; // [nT*3]
    int nT;
    int tris_per_chunk; // multiple of kExactTile
    int nchunks;
    const float* q;
    int64_t n;
    GridDesc g; // grid mode: n = nx*ny*(z1-z0), x fastest
    float* partial; // [nchunks][n] (nchunks > 1) else unused
    float* out_omega;
    uint8_t* out_inside;
};

template <bool GRID>
__device__ __forceinline__ void exact_query_point(const ExactArgs& a, int64_t i, float& x, float& y, float& z)
{
    if (GRID) {
        const int ix = (int)(i % a.g.nx);
        const int iy = (int)((i / a.g.nx) % a.g.ny);
        const int iz = (int)(i / ((int64_t)a.g.nx * a.g.ny)) + a.g.z0;
        x = wn_lattice_coord(a.g.ox, a.g.sx, ix);
        y = wn_lattice_coord(a.g.oy, a.g.sy, iy);
        z = wn_lattice_coord(a.g.oz, a.g.sz, iz);
    } else {
        x = __ldg(a.q + 3 * i);
        y = __ldg(a.q + 3 * i + 1);
        z = __ldg(a.q + 3 * i + 2);
    }
}

template <bool GRID>
__global__ void __launch_bounds__(256) k_exact(const ExactArgs a)
{
    __shared__ float4 sh[kExactTile * 3];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < a.n;
    float x = 0, y = 0, z = 0;
    if (valid) exact_query_point<GRID>(a, i, x, y, z);
    const int t_begin = blockIdx.y * a.tris_per_chunk;
    const int t_end = min(a.nT, t_begin + a.tris_per_chunk);
    float sum = 0.0f, comp = 0.0f; // Kahan over tile sums
    for (int t0 = t_begin; t0 < t_end; t0 += kExactTile) {
        const int cnt = min(kExactTile, t_end - t0);
        __syncthreads();
        for (int j = threadIdx.x; j < cnt * 3; j += blockDim.x) sh[j] = __ldg(a.tris + 3 * (int64_t)t0 + j);
        __syncthreads();
        float tile = 0.0f;
#pragma unroll 4
        for (int t = 0; t < cnt; ++t) tile += wn_tri_solid_angle(x, y, z, sh[3 * t], sh[3 * t + 1], sh[3 * t + 2]);
        const float yk = tile - comp;
        const float tk = sum + yk;
        comp = (tk - sum) - yk;
        sum = tk;
    }
    if (!valid) return;
    if (a.nchunks > 1) {
        a.partial[(int64_t)blockIdx.y * a.n + i] = sum;
    } else {
        if (a.out_omega) a.out_omega[i] = sum;
        if (a.out_inside) a.out_inside[i] = wn_inside_from_omega(sum) ? 1 : 0;
    }
}

// small batches: one warp per (query, chunk); lanes stride over the chunk's triangles, shuffle reduction
template <bool GRID>
__global__ void __launch_bounds__(256) k_exact_small(const ExactArgs a)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + wid;
    if (i >= a.n) return;
    float x, y, z;
    exact_query_point<GRID>(a, i, x, y, z);
    const int t_begin = blockIdx.y * a.tris_per_chunk;
    const int t_end = min(a.nT, t_begin + a.tris_per_chunk);
    float sum = 0.0f;
    for (int t = t_begin + lane; t < t_end; t += 32)
        sum += wn_tri_solid_angle(x, y, z, __ldg(a.tris + 3 * (int64_t)t), __ldg(a.tris + 3 * (int64_t)t + 1), __ldg(a.tris + 3 * (int64_t)t + 2));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
    if (lane == 0) {
        if (a.nchunks > 1) {
            a.partial[(int64_t)blockIdx.y * a.n + i] = sum;
        } else {
            if (a.out_omega) a.out_omega[i] = sum;
            if (a.out_inside) a.out_inside[i] = wn_inside_from_omega(sum) ? 1 : 0;
        }
    }
}

__global__ void __launch_bounds__(256) k_exact_reduce(const float* __restrict__ partial, int nchunks, int64_t n, float* __restrict__ out_omega,
                                                      uint8_t* __restrict__ out_inside)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float sum = 0.0f, comp = 0.0f;
    for (int c = 0; c < nchunks; ++c) {
        const float yk = partial[(int64_t)c * n + i] - comp;
        const float tk = sum + yk;
        comp = (tk - sum) - yk;
        sum = tk;
    }
    if (out_omega) out_omega[i] = sum;
    if (out_inside) out_inside[i] = wn_inside_from_omega(sum) ? 1 : 0;
}

// ---- K9 Morton keys of query points ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_point_bounds(const float* __restrict__ q, int64_t n, int* __restrict__ bounds)
{
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float c = __ldg(q + 3 * i + k);
            if (fabsf(c) <= 3.402823466e38f) { // ignore NaN / inf
                lo[k] = fminf(lo[k], c);
                hi[k] = fmaxf(hi[k], c);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(kFull, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(kFull, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&bounds[k], float_to_ordered(lo[k]));
            atomicMax(&bounds[3 + k], float_to_ordered(hi[k]));
        }
    }
}

__global__ void __launch_bounds__(256) k_point_morton(const float* __restrict__ q, int64_t n, const int* __restrict__ bounds,
                                                      uint32_t* __restrict__ keys, unsigned* __restrict__ vals)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float lx = ordered_to_float(bounds[0]), ly = ordered_to_float(bounds[1]), lz = ordered_to_float(bounds[2]);
    const float ex = ordered_to_float(bounds[3]) - lx, ey = ordered_to_float(bounds[4]) - ly, ez = ordered_to_float(bounds[5]) - lz;
    const float ext = fmaxf(ex, fmaxf(ey, ez));
    const float inv = ext > 0.0f ? 1.0f / ext : 0.0f;
    const float x = (__ldg(q + 3 * i) - lx) * inv, y = (__ldg(q + 3 * i + 1) - ly) * inv, z = (__ldg(q + 3 * i + 2) - lz) * inv;
    keys[i] = (uint32_t)wn_morton(x, y, z, 10);
    vals[i] = (unsigned)i;
}

// ---- FP32 FMA peak probe (roofline denominator for the FMA-bound kernels; MEASURED_PEAKS.json has no FP32 figure) --
__global__ void __launch_bounds__(256) k_fma_peak(int iters, float* __restrict__ sink)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-3f * blockIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fmaf(a0, m, c);
            a1 = fmaf(a1, m, c);
            a2 = fmaf(a2, m, c);
            a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c);
            a5 = fmaf(a5, m, c);
            a6 = fmaf(a6, m, c);
            a7 = fmaf(a7, m, c);
        }
    }
    const float s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678f) sink[0] = s; // never true; keeps the chain alive
}

} // namespace wn
