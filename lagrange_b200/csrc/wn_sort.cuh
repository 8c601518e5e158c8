// wn_sort.cuh — K2 / K9: stable LSD radix sort of (key, value) pairs, 8-bit digits, hand-written for sm_100a.
//
// Used for (a) the triangle Morton codes of the LBVH build (u64 keys, 63 significant bits) and (b) the Morton codes
// of incoherent query sets (u32 keys, 30 bits). Per 8-bit pass:
//   k_radix_hist    per-tile digit histogram (smem atomics)                        -> hist[digit][tile]
//   wn_scan_u32     exclusive scan of the digit-major table (3 small kernels)      -> global base of (digit, tile)
//   k_radix_scatter re-reads the tile, ranks every item among equal digits with warp match_any + per-warp smem
//                   counters (stable: tile order = warp-major, round, lane), scatters keys and values
// HBM traffic per pass: read keys twice + values once, write both once: (2*K + V + K + V) bytes per item.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace wn {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16; // per thread
constexpr int kSortTile = kSortThreads * kSortItems;

template <typename K>
__device__ __forceinline__ uint32_t radix_digit(K key, int shift, uint32_t mask)
{
    return (uint32_t)(key >> shift) & mask;
}

template <typename K>
__global__ void __launch_bounds__(kSortThreads) k_radix_hist(const K* __restrict__ keys, int64_t n, int shift, uint32_t mask,
                                                             int ntiles, uint32_t* __restrict__ hist /* [256][ntiles] */)
{
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int64_t idx = base + (int64_t)r * kSortThreads + threadIdx.x;
        if (idx < n) atomicAdd(&sh[radix_digit(keys[idx], shift, mask)], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] = sh[threadIdx.x];
}

// ---- generic exclusive scan of a u32 array (in place), n up to 2^31 ---------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* sh_warp /* 32 */, uint32_t& block_total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (blockDim.x >> 5) ? sh_warp[lane] : 0;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        sh_warp[lane] = winc - w; // exclusive warp offsets
        if (lane == 31) sh_warp[32] = winc;
    }
    __syncthreads();
    block_total = sh_warp[32];
    const uint32_t r = sh_warp[warp] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const uint32_t* __restrict__ data, int64_t n, uint32_t* __restrict__ tile_sums)
{
    __shared__ uint32_t sh[33];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) s += data[base + k];
    uint32_t total;
    block_exclusive_scan(s, sh, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of m tile sums in place (m is small: n / 2048)
__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(uint32_t* __restrict__ tile_sums, int m)
{
    __shared__ uint32_t sh[33];
    uint32_t carry = 0;
    for (int base = 0; base < m; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < m ? tile_sums[i] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, sh, total);
        if (i < m) tile_sums[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(uint32_t* __restrict__ data, int64_t n, const uint32_t* __restrict__ tile_sums)
{
    __shared__ uint32_t sh[33];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = base + k < n ? data[base + k] : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan(s, sh, total) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
}

// scratch: at least scan_scratch_elems(n) uint32
inline int64_t scan_scratch_elems(int64_t n)
{
    return (n + kScanTile - 1) / kScanTile + 1;
}

inline void exclusive_scan_u32(uint32_t* data, int64_t n, uint32_t* scratch, cudaStream_t stream)
{
    if (n <= 0) return;
    const int m = (int)((n + kScanTile - 1) / kScanTile);
    k_scan_reduce<<<m, kScanThreads, 0, stream>>>(data, n, scratch);
    k_scan_tiles<<<1, kScanThreads, 0, stream>>>(scratch, m);
    k_scan_apply<<<m, kScanThreads, 0, stream>>>(data, n, scratch);
}

// ---- scatter ---------------------------------------------------------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(kSortThreads) k_radix_scatter(const K* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                K* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n,
                                                                int shift, uint32_t mask, int ntiles,
                                                                const uint32_t* __restrict__ offs /* scanned hist */)
{
    __shared__ uint32_t wcount[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&wcount[0][0])[i] = 0;
    __syncthreads();

    // tile order: warp-major, then round, then lane  (each warp owns a contiguous 32*kSortItems slice of the tile)
    const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * (32 * kSortItems);
    K key[kSortItems];
    uint32_t rank[kSortItems];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int64_t idx = wbase + r * 32 + lane;
        const bool valid = idx < n;
        key[r] = valid ? keys_in[idx] : (K)0;
        const uint32_t d = valid ? radix_digit(key[r], shift, mask) : 256u; // 256 = sentinel group for the tail
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader && valid) {
            old = wcount[warp][d];
            wcount[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over warps per digit, seeded with the global base of (digit, tile)
    {
        const int d = threadIdx.x; // 256 threads, one digit each
        uint32_t run = offs[(int64_t)d * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = wcount[w][d];
            wcount[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int64_t idx = wbase + r * 32 + lane;
        if (idx < n) {
            const uint32_t d = radix_digit(key[r], shift, mask);
            const uint32_t pos = wcount[warp][d] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = vals_in[idx];
        }
    }
}

struct SortScratch
{
    uint32_t* hist = nullptr; // 256 * ntiles
    uint32_t* scan = nullptr; // scan_scratch_elems(256 * ntiles)
};

inline int64_t sort_ntiles(int64_t n)
{
    return (n + kSortTile - 1) / kSortTile;
}
inline int64_t sort_scratch_bytes(int64_t n)
{
    const int64_t nt = sort_ntiles(n);
    return (256 * nt + scan_scratch_elems(256 * nt)) * (int64_t)sizeof(uint32_t);
}

// Sorts bits [begin_bit, end_bit) of the keys. Ping-pongs between (k0,v0) and (k1,v1); returns 0 if the result is in
// (k0,v0), 1 if in (k1,v1). `scratch` must hold sort_scratch_bytes(n).
template <typename K>
inline int radix_sort_pairs(K* k0, uint32_t* v0, K* k1, uint32_t* v1, int64_t n, int begin_bit, int end_bit, void* scratch,
                            cudaStream_t stream)
{
    if (n <= 1) return 0;
    const int ntiles = (int)sort_ntiles(n);
    uint32_t* hist = (uint32_t*)scratch;
    uint32_t* scan = hist + (int64_t)256 * ntiles;
    int cur = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        K* kin = cur ? k1 : k0;
        K* kout = cur ? k0 : k1;
        uint32_t* vin = cur ? v1 : v0;
        uint32_t* vout = cur ? v0 : v1;
        const int nbits = end_bit - shift < 8 ? end_bit - shift : 8;
        const uint32_t mask = (1u << nbits) - 1u;
        k_radix_hist<K><<<ntiles, kSortThreads, 0, stream>>>(kin, n, shift, mask, ntiles, hist);
        exclusive_scan_u32(hist, (int64_t)256 * ntiles, scan, stream);
        k_radix_scatter<K><<<ntiles, kSortThreads, 0, stream>>>(kin, vin, kout, vout, n, shift, mask, ntiles, hist);
        cur ^= 1;
    }
    return cur;
}

} // namespace wn
