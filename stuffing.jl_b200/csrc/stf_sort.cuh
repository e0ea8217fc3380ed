// stf_sort.cuh — device radix sort (stable, LSD, 8-bit digits) and exclusive scan.
// Replaces the Dict / linked-list bucketing of HashSpacialQTree and LinkedSpacialQTree
// (qtrees.jl:411-415, 501-538) and the final sort! of the many-body calls (qtree_functions.jl:252).
//
// Granularity: one WARP owns a contiguous slice of the input and a private 256-bin histogram in
// shared memory.  Ranks inside a 32-key chunk come from __match_any_sync, so equal digits keep
// their input order (stable).  Three launches per pass: histogram, scan over (digit, warp), scatter.
#pragma once
#include "stf_common.cuh"

#define SORT_WARPS_PER_BLOCK 8
#define SORT_THREADS (32 * SORT_WARPS_PER_BLOCK)

struct SortPlan { int nvw; int per_warp; int blocks; };
static inline SortPlan sort_plan(int64_t n) {
    SortPlan p;
    int64_t nvw = (n + 511) / 512;            // >= 512 keys per warp
    if (nvw < 1) nvw = 1;
    if (nvw > 148 * 16) nvw = 148 * 16;
    nvw = (nvw + SORT_WARPS_PER_BLOCK - 1) / SORT_WARPS_PER_BLOCK * SORT_WARPS_PER_BLOCK;
    int64_t per = (n + nvw - 1) / nvw;
    per = (per + 31) / 32 * 32;
    p.nvw = (int)nvw; p.per_warp = (int)(per < 32 ? 32 : per); p.blocks = (int)(nvw / SORT_WARPS_PER_BLOCK);
    return p;
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const uint64_t* __restrict__ keys, int64_t n, int shift, int per_warp, int nvw, uint32_t* __restrict__ hist) {
    __shared__ uint32_t cnt[SORT_WARPS_PER_BLOCK][256];
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int vw = blockIdx.x * SORT_WARPS_PER_BLOCK + w;
    for (int d = lane; d < 256; d += 32) cnt[w][d] = 0;
    __syncwarp();
    int64_t beg = (int64_t)vw * per_warp, end = min(n, beg + per_warp);
    for (int64_t i = beg + lane; i < end; i += 32) atomicAdd(&cnt[w][(keys[i] >> shift) & 255], 1u);
    __syncwarp();
    for (int d = lane; d < 256; d += 32) hist[(size_t)d * nvw + vw] = cnt[w][d];
}

// exclusive scan of `m` uint32 by ONE block (m = 256*nvw <= ~600k; a few microseconds)
__global__ void __launch_bounds__(1024) k_scan_u32(uint32_t* __restrict__ data, int64_t m, uint32_t* __restrict__ total) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry, btot;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < m; base += 4096) {
        int64_t i0 = base + (int64_t)threadIdx.x * 4;
        uint32_t v[4]; uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { v[k] = (i0 + k < m) ? data[i0 + k] : 0; s += v[k]; }
        uint32_t incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        if (w == 0) {
            uint32_t x = wsum[lane], xi = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += t; }
            wsum[lane] = xi - x; // exclusive warp offsets
            if (lane == 31) btot = xi;
        }
        __syncthreads();
        uint32_t run = carry + wsum[w] + (incl - s);
#pragma unroll
        for (int k = 0; k < 4; k++) { if (i0 + k < m) data[i0 + k] = run; run += v[k]; }
        __syncthreads();
        if (threadIdx.x == 0) carry += btot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n, int shift,
                                                               int per_warp, int nvw, const uint32_t* __restrict__ hist,
                                                               uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t base[SORT_WARPS_PER_BLOCK][256];
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int vw = blockIdx.x * SORT_WARPS_PER_BLOCK + w;
    for (int d = lane; d < 256; d += 32) base[w][d] = hist[(size_t)d * nvw + vw];
    __syncwarp();
    int64_t beg = (int64_t)vw * per_warp, end = min(n, beg + per_warp);
    for (int64_t i0 = beg; i0 < end; i0 += 32) {
        int64_t i = i0 + lane;
        bool valid = i < end;
        uint64_t k = valid ? keys[i] : 0;
        uint32_t v = valid ? vals[i] : 0;
        uint32_t d = valid ? (uint32_t)((k >> shift) & 255) : 256u + lane; // invalid lanes never match
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t rank = __popc(peers & ((1u << lane) - 1));
        uint32_t dst = 0;
        if (valid) dst = base[w][d] + rank;
        __syncwarp();
        if (valid && rank == 0) base[w][d] += __popc(peers);
        __syncwarp();
        if (valid) { keys_out[dst] = k; vals_out[dst] = v; }
    }
}

// Sorts (keys, vals) by bits [bit_lo, bit_hi) of the key.  Buffers ping-pong; returns 0 if the result
// is in (keys, vals), 1 if it is in (keys_alt, vals_alt).  `hist` holds >= 256*nvw uint32.
static inline int radix_sort_pairs(cudaStream_t st, int64_t* launches, uint64_t* keys, uint32_t* vals, uint64_t* keys_alt, uint32_t* vals_alt,
                                   int64_t n, int bit_lo, int bit_hi, uint32_t* hist) {
    if (n <= 1) return 0;
    SortPlan p = sort_plan(n);
    int flip = 0;
    for (int shift = bit_lo; shift < bit_hi; shift += 8) {
        uint64_t* kin = flip ? keys_alt : keys; uint32_t* vin = flip ? vals_alt : vals;
        uint64_t* kout = flip ? keys : keys_alt; uint32_t* vout = flip ? vals : vals_alt;
        k_sort_hist<<<p.blocks, SORT_THREADS, 0, st>>>(kin, n, shift, p.per_warp, p.nvw, hist);
        k_scan_u32<<<1, 1024, 0, st>>>(hist, (int64_t)256 * p.nvw, nullptr);
        k_sort_scatter<<<p.blocks, SORT_THREADS, 0, st>>>(kin, vin, n, shift, p.per_warp, p.nvw, hist, kout, vout);
        flip ^= 1; if (launches) *launches += 3;
    }
    return flip;
}
static inline size_t radix_sort_hist_words(int64_t n) { return (size_t)256 * sort_plan(n).nvw; }

// first INVALID key of a sorted array (multi-launch path), so that both sort paths publish the valid count
__global__ void k_count_valid(const uint64_t* __restrict__ keys, int64_t n, unsigned long long* __restrict__ n_out) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (keys[mid] != STF_KEY_INVALID) lo = mid + 1; else hi = mid; }
    *n_out = (unsigned long long)lo;
}
