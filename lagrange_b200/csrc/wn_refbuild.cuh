// wn_refbuild.cuh — K3R on sm_100a: the kernels behind wn_refbuild_core.cuh's driver (WN_HIERARCHY_REFERENCE).
// Replaces the hierarchy half of UT_SolidAngle::init (adobe/lagrange modules/winding/src/FastWindingNumber.cpp:57).
// All passes are HBM/latency-bound integer + min/max work: one thread per item or per node, coalesced arrays, grids sized to
// the element count; the only contended atomics (span boxes of the few large ranges near the root) are privatised per CTA in
// shared memory.
#pragma once

#include <cuda_runtime.h>

#include "wn_build.cuh"
#include "wn_refbuild_core.cuh"

namespace wn {

template <class F>
__global__ void __launch_bounds__(kBuildThreads) k_ref_for_each(const F f, const int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f(i);
}

constexpr int kRefBinItems = 8; // positions per thread of k_ref_bin

// WnRefBin with the span table of the CTA's first range kept in shared memory: near the root one range covers whole CTAs,
// and its 16 x 7 words would otherwise take every item's atomics.
__global__ void __launch_bounds__(kBuildThreads) k_ref_bin(const WnRefState s)
{
    __shared__ int sh[WN_REF_ROW];
    __shared__ int sh_row;
    const int64_t base = (int64_t)blockIdx.x * (kBuildThreads * kRefBinItems);
    if (threadIdx.x == 0) {
        int row = -1;
        const int i = s.owner[base];
        if (i >= 0) {
            const WnRefTask t = s.tasks[i];
            if (t.mode == 1 && base >= t.ts && base < t.ts + t.tn) row = t.ts / WN_REF_MID;
        }
        sh_row = row;
    }
    for (int k = threadIdx.x; k < WN_REF_ROW; k += kBuildThreads) sh[k] = k >= WN_REF_NSPANS * 6 ? 0 : ((k % 6) < 3 ? 0x7fffffff : (int)0x80000000);
    __syncthreads();
    const int my_row = sh_row;
    for (int k = 0; k < kRefBinItems; ++k) {
        const int64_t p = base + (int64_t)k * kBuildThreads + threadIdx.x;
        if (p >= s.N) break;
        const int i = s.owner[p];
        if (i < 0) continue;
        const WnRefTask t = s.tasks[i];
        if (t.mode != 1 || p < t.ts || p >= t.ts + t.tn) continue;
        float b[6];
        wn_ref_tri_box(s.tbox, s.idx[p], b);
        const int sp = wn_ref_span(t, b);
        const int row_id = t.ts / WN_REF_MID;
        int* row = row_id == my_row ? sh : s.rows + (size_t)row_id * WN_REF_ROW;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&row[sp * 6 + a], wn_ref_ordered(b[a]));
            atomicMax(&row[sp * 6 + 3 + a], wn_ref_ordered(b[3 + a]));
        }
        atomicAdd(&row[WN_REF_NSPANS * 6 + sp], 1);
    }
    __syncthreads();
    if (my_row >= 0) {
        int* row = s.rows + (size_t)my_row * WN_REF_ROW;
        for (int k = threadIdx.x; k < WN_REF_ROW; k += kBuildThreads) {
            const int v = sh[k];
            if (k >= WN_REF_NSPANS * 6) {
                if (v) atomicAdd(&row[k], v);
            } else if ((k % 6) < 3) {
                if (v != 0x7fffffff) atomicMin(&row[k], v);
            } else if (v != (int)0x80000000) {
                atomicMax(&row[k], v);
            }
        }
    }
}

// union of all triangle boxes -> rows[0..5] (ordered ints); rows[0..5] preset to the empty box by the launcher
__global__ void __launch_bounds__(kBuildThreads) k_ref_root_bounds(const WnRefState s)
{
    int enc[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < s.N; t += (int64_t)gridDim.x * blockDim.x) {
        float b[6];
        wn_ref_tri_box(s.tbox, (unsigned)t, b);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            enc[a] = min(enc[a], wn_ref_ordered(b[a]));
            enc[3 + a] = max(enc[3 + a], wn_ref_ordered(b[3 + a]));
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        enc[a] = __reduce_min_sync(0xffffffffu, enc[a]);
        enc[3 + a] = __reduce_max_sync(0xffffffffu, enc[3 + a]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&s.rows[a], enc[a]);
            atomicMax(&s.rows[3 + a], enc[3 + a]);
        }
    }
}

struct RefCudaBackend
{
    cudaStream_t st = nullptr;
    uint32_t* scan_scratch = nullptr;
    void* sort_scratch = nullptr;
    uint64_t* keys_alt = nullptr;
    cudaError_t err = cudaSuccess;
    int syncs = 0;

    void note(cudaError_t e)
    {
        if (err == cudaSuccess && e != cudaSuccess) err = e;
    }
    template <class F>
    void for_each(int64_t n, const F& f)
    {
        if (n <= 0) return;
        k_ref_for_each<F><<<grid_for(n), kBuildThreads, 0, st>>>(f, n);
    }
    void bin(const WnRefState& s)
    {
        const int64_t per = kBuildThreads * kRefBinItems;
        k_ref_bin<<<(int)((s.N + per - 1) / per), kBuildThreads, 0, st>>>(s);
    }
    void root_bounds(const WnRefState& s)
    {
        const int init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
        write(s.rows, init, sizeof(init));
        k_ref_root_bounds<<<std::min(grid_for(s.N), 148 * 8), kBuildThreads, 0, st>>>(s);
    }
    void scan(uint32_t* d, int64_t n) { exclusive_scan_u32(d, n, scan_scratch, st); }
    int sort64(uint64_t* keys, unsigned* vals, unsigned* vals_alt, int64_t n, int end_bit)
    {
        return radix_sort_pairs<uint64_t>(keys, vals, keys_alt, vals_alt, n, 0, end_bit, sort_scratch, st);
    }
    void read(int* h, const int* d, int n)
    {
        note(cudaMemcpyAsync(h, d, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
        note(cudaStreamSynchronize(st));
        ++syncs;
    }
    void zero(void* p, size_t bytes) { note(cudaMemsetAsync(p, 0, bytes, st)); }
    void write(void* d, const void* h, size_t bytes) { note(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st)); } // pageable: staged before it returns
};

} // namespace wn
