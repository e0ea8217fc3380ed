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

// =============================================================================================
// single-launch sort for latency-bound sizes (n <= SORT_SMALL_MAX): ONE CTA runs
//   phase 0  ordered compaction of the valid records (key != INVALID), and the OR of all key
//            differences, from which the digit windows are chosen on the device — bit ranges in which no
//            two keys differ are never sorted;
//   phase 1+ stable 8-bit LSD passes (warp-private histograms, block scan over (digit, warp),
//            __match_any_sync ranks);
// and leaves the result in (keys, vals) whatever the number of passes.  n may live on the device.
// If the valid records fit (n <= SS_TINY_CAP and key < 2^(64-vbits)), (key,val) are packed into one
// 64-bit word and ALL passes run in shared memory: one global read and one global write in total.
// =============================================================================================
#define SS_THREADS 1024
#define SS_WARPS 32
#define SS_UNROLL 4
#define SORT_SMALL_MAX (1 << 18)
#define SS_TINY_CAP 11264
#define SS_TINY_SMEM (2 * 8 * SS_TINY_CAP)

__device__ __forceinline__ uint32_t ss_block_exclusive(uint32_t v, uint32_t* wsum, uint32_t* total) {
    // exclusive scan of one value per thread over the block; all threads must call
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t x = wsum[lane], xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += t; }
        wsum[lane] = xi - x;
        if (lane == 31) *total = xi;
    }
    __syncthreads();
    uint32_t r = wsum[w] + incl - v;
    return r;
}

// one stable 8-bit pass over n elements.  PACKED: one u64 array (key<<vbits | val); else split (key, val) arrays
__device__ long long g_sort_dbg[16];
#define SS_STAMP(i) do { if (threadIdx.x == 0) g_sort_dbg[i] = clock64() - t_start; } while (0)
template <bool PACKED>
__device__ __forceinline__ void ss_pass(const uint64_t* __restrict__ ksrc, const uint32_t* __restrict__ vsrc, uint64_t* __restrict__ kdst, uint32_t* __restrict__ vdst,
                                        int64_t n, int shift, uint32_t (*cnt)[256], uint32_t* wsum, uint32_t* s_total) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int64_t per = ((n + SS_WARPS - 1) / SS_WARPS + 31) / 32 * 32;
    int64_t beg = (int64_t)w * per, end = min(n, beg + per);
    long long tp0 = clock64();
    for (int d = lane; d < 256; d += 32) cnt[w][d] = 0;
    __syncwarp();
    for (int64_t i0 = beg; i0 < end; i0 += 32 * SS_UNROLL) {
        uint64_t k[SS_UNROLL];
#pragma unroll
        for (int u = 0; u < SS_UNROLL; u++) { int64_t i = i0 + 32 * u + lane; k[u] = i < end ? ksrc[i] : 0; }
#pragma unroll
        for (int u = 0; u < SS_UNROLL; u++) { int64_t i = i0 + 32 * u + lane; if (i < end) atomicAdd(&cnt[w][(k[u] >> shift) & 255], 1u); }
    }
    __syncthreads();
    long long tp1 = clock64();
    { // exclusive scan over (digit, warp): thread t owns digit t>>2, warps (t&3)*8 .. +8
        int d = threadIdx.x >> 2, w0 = (threadIdx.x & 3) * 8;
        uint32_t loc[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { loc[j] = cnt[w0 + j][d]; sum += loc[j]; }
        uint32_t run = ss_block_exclusive(sum, wsum, s_total);
#pragma unroll
        for (int j = 0; j < 8; j++) { cnt[w0 + j][d] = run; run += loc[j]; }
    }
    __syncthreads();
    long long tp2 = clock64();
    for (int64_t i0 = beg; i0 < end; i0 += 32 * SS_UNROLL) {
        uint64_t k[SS_UNROLL]; uint32_t v[SS_UNROLL];
#pragma unroll
        for (int u = 0; u < SS_UNROLL; u++) {
            int64_t i = i0 + 32 * u + lane; bool in = i < end;
            k[u] = in ? ksrc[i] : 0; v[u] = (!PACKED && in) ? vsrc[i] : 0;
        }
#pragma unroll
        for (int u = 0; u < SS_UNROLL; u++) {
            int64_t i = i0 + 32 * u + lane;
            bool valid = i < end;
            uint32_t d = valid ? (uint32_t)((k[u] >> shift) & 255) : 256u + lane;
            uint32_t peers = __match_any_sync(0xffffffffu, d);
            uint32_t rank = __popc(peers & ((1u << lane) - 1));
            uint32_t dst = valid ? cnt[w][d] + rank : 0;
            __syncwarp();
            if (valid && rank == 0) cnt[w][d] += __popc(peers);
            __syncwarp();
            if (valid) { kdst[dst] = k[u]; if (!PACKED) vdst[dst] = v[u]; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { long long tp3 = clock64(); g_sort_dbg[4] = tp1 - tp0; g_sort_dbg[5] = tp2 - tp1; g_sort_dbg[6] = tp3 - tp2; }
}

__global__ void __launch_bounds__(SS_THREADS) k_sort_small(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint64_t* __restrict__ keys_alt,
                                                           uint32_t* __restrict__ vals_alt, const unsigned long long* __restrict__ n_dev, int64_t n_host,
                                                           int compact_invalid, int max_bit, int vbits, unsigned long long* __restrict__ n_out) {
    extern __shared__ uint64_t tiny[]; // 2 x SS_TINY_CAP packed words (only when launched with SS_TINY_SMEM)
    __shared__ uint32_t cnt[SS_WARPS][256];
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t s_total;
    __shared__ unsigned long long s_or[SS_WARPS];
    __shared__ int s_shift[9], s_npass;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long t_start = clock64();
    int64_t n = n_dev ? (int64_t)*n_dev : n_host;
    if (n > n_host) n = n_host;
    bool can_pack = vbits > 0 && max_bit + vbits <= 64;
    uint64_t* ksrc = keys; uint32_t* vsrc = vals; uint64_t* kdst = keys_alt; uint32_t* vdst = vals_alt;

    // ---- phase 0: stable compaction → alt buffers (and, packed, into shared memory while it fits)
    uint32_t base = 0;
    for (int64_t c0 = 0; c0 < n; c0 += SS_THREADS * SS_UNROLL) {
        uint64_t k[SS_UNROLL]; uint32_t v[SS_UNROLL]; uint32_t nv = 0;
        int64_t i0 = c0 + (int64_t)threadIdx.x * SS_UNROLL; // contiguous per thread keeps the order
#pragma unroll
        for (int u = 0; u < SS_UNROLL; u++) {
            bool in = i0 + u < n;
            k[u] = in ? ksrc[i0 + u] : STF_KEY_INVALID; v[u] = in ? vsrc[i0 + u] : 0;
            nv += (in && !(compact_invalid && k[u] == STF_KEY_INVALID));
        }
        uint32_t off = ss_block_exclusive(nv, wsum, &s_total) + base;
#pragma unroll
        for (int u = 0; u < SS_UNROLL; u++) {
            bool in = i0 + u < n;
            if (in && !(compact_invalid && k[u] == STF_KEY_INVALID)) {
                kdst[off] = k[u]; vdst[off] = v[u];
                if (can_pack && off < SS_TINY_CAP) tiny[off] = (k[u] << vbits) | v[u];
                off++;
            }
        }
        base += s_total;
        __syncthreads();
    }
    SS_STAMP(0);
    n = base;
    { uint64_t* tk = ksrc; ksrc = kdst; kdst = tk; uint32_t* tv = vsrc; vsrc = vdst; vdst = tv; }
    bool packed = can_pack && n <= SS_TINY_CAP;
    __syncthreads();

    // ---- digit windows: only bit ranges in which keys differ
    unsigned long long vary = 0;
    if (packed) { uint64_t p0 = n > 0 ? tiny[0] : 0; for (int64_t i = threadIdx.x; i < n; i += SS_THREADS) vary |= tiny[i] ^ p0; vary >>= vbits; }
    else { uint64_t key0 = n > 0 ? ksrc[0] : 0; for (int64_t i = threadIdx.x; i < n; i += SS_THREADS) vary |= ksrc[i] ^ key0; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vary |= __shfl_xor_sync(0xffffffffu, vary, o);
    if (lane == 0) s_or[w] = vary;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long V = 0;
        for (int i = 0; i < SS_WARPS; i++) V |= s_or[i];
        if (max_bit < 64) V &= (1ull << max_bit) - 1;
        int np = 0;
        while (V) { int p = __ffsll((long long)V) - 1; s_shift[np++] = p; V &= (p + 8 >= 64) ? 0ull : ~((1ull << (p + 8)) - 1); }
        s_npass = np;
        if (n_out) *n_out = (unsigned long long)n;
    }
    __syncthreads();
    int npass = s_npass;
    SS_STAMP(1);
    if (threadIdx.x == 0) { g_sort_dbg[8] = n; g_sort_dbg[9] = npass; g_sort_dbg[10] = packed; }

    if (packed) {
        uint64_t* a = tiny; uint64_t* b = tiny + SS_TINY_CAP;
        for (int pass = 0; pass < npass; pass++) {
            ss_pass<true>(a, nullptr, b, nullptr, n, s_shift[pass] + vbits, cnt, wsum, &s_total);
            uint64_t* t = a; a = b; b = t;
        }
        SS_STAMP(2);
        uint64_t vmask = (1ull << vbits) - 1;
        for (int64_t i = threadIdx.x; i < n; i += SS_THREADS) { uint64_t p = a[i]; keys[i] = p >> vbits; vals[i] = (uint32_t)(p & vmask); }
        SS_STAMP(3);
        return;
    }
    for (int pass = 0; pass < npass; pass++) {
        ss_pass<false>(ksrc, vsrc, kdst, vdst, n, s_shift[pass], cnt, wsum, &s_total);
        uint64_t* tk = ksrc; ksrc = kdst; kdst = tk; uint32_t* tv = vsrc; vsrc = vdst; vdst = tv;
    }
    if (ksrc != keys) { // an odd number of hops: bring the result home
        for (int64_t i = threadIdx.x; i < n; i += SS_THREADS) { keys[i] = ksrc[i]; vals[i] = vsrc[i]; }
    }
}
// first INVALID key of a sorted array (multi-launch path), so that both sort paths publish the valid count
__global__ void k_count_valid(const uint64_t* __restrict__ keys, int64_t n, unsigned long long* __restrict__ n_out) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (keys[mid] != STF_KEY_INVALID) lo = mid + 1; else hi = mid; }
    *n_out = (unsigned long long)lo;
}
