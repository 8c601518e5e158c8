// wn_exact.cuh — K7 exact all-pairs mode, K9 query Morton keys, FP32 FMA peak probe (sm_100a).
//
// K7 k_exact : exact mode (config 5, on-device ground truth): sum of UTsignedSolidAngleTri over ALL triangles (SURVEY.md
//              A.1). Triangles staged in shared memory tiles, one query per thread, tile partial sums added with
//              compensated summation; k_exact_small splits the triangles of one query over a warp and reduces with
//              shuffles (small batches). Partial sums over triangle chunks are combined by k_exact_reduce.
//              CUDA-core FMA + MUFU (rsqrt, atan) bound: 75 flop per (query, triangle) pair; no tensor cores by design.
// K9 k_point_bounds / k_point_morton : 30-bit Morton keys of incoherent query sets for wn::radix_sort_pairs.
#pragma once

#include <cuda_runtime.h>

#include "wn_device.cuh"
#include "wn_build.cuh"

namespace wn {

constexpr unsigned kFull = 0xffffffffu;

struct GridDesc
{
    float ox, oy, oz, sx, sy, sz;
    int nx, ny, nz;
    int z0, z1; // slab
};

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

// Sum of the solid angles of `cnt` triangles (3 float4 each, shared memory) seen from (x, y, z). Groups of WN_EXACT_GROUP
// triangles go through one complex product and one atan2 (wn_tri_fold); a group with a triangle that subtends a large angle from
// this point (or touches it) is redone term by term with the reference formulation and its zero rules.
__device__ __forceinline__ float exact_tile_sum(float x, float y, float z, const float4* __restrict__ sh, int cnt)
{
    float tile = 0.0f;
    int t = 0;
    for (; t + WN_EXACT_GROUP <= cnt; t += WN_EXACT_GROUP) {
        float zr = 1.0f, zi = 0.0f;
        bool ok = true;
#pragma unroll
        for (int j = 0; j < WN_EXACT_GROUP; ++j) ok = wn_tri_fold(x, y, z, sh[3 * (t + j)], sh[3 * (t + j) + 1], sh[3 * (t + j) + 2], zr, zi) && ok;
        if (ok) {
            tile += 2.0f * atan2f(zi, zr);
        } else {
            for (int j = 0; j < WN_EXACT_GROUP; ++j) tile += wn_tri_solid_angle(x, y, z, sh[3 * (t + j)], sh[3 * (t + j) + 1], sh[3 * (t + j) + 2]);
        }
    }
    for (; t < cnt; ++t) tile += wn_tri_solid_angle(x, y, z, sh[3 * t], sh[3 * t + 1], sh[3 * t + 2]);
    return tile;
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
        float tile = exact_tile_sum(x, y, z, sh, cnt);
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
