// stf_blocksort.cuh — latency-oriented LSD radix sort of one CTA (1024 threads) over <= BS_CAP packed 64-bit words.
//
// Single-scene rounds (10k objects, a few thousand hits) are latency-bound: a multi-launch radix sort costs three
// launches per digit for half a megabyte of data.  Here the whole sort is a device function that fused kernels call:
// the elements live in REGISTERS between passes (element e = w*per + c*32 + lane of warp w, chunk c), ranks inside a
// 32-element chunk come from eight ballots (no match.any, whose latency grows with the number of distinct digits),
// every warp owns a private 256-bin histogram row (padded to 257 words: the (digit, warp) scan is bank-conflict free),
// and only bit windows in which the keys actually differ are sorted.  Stable.
// Measured on B200 (cycles for n = 2833 / 5666): this variant 28k / 22k; 5-bit digits with per-thread 16-bit counters
// 40k / 21k; block merge sort (merge path, 10 levels) 47k / 75k — one SM's shared-memory bandwidth and issue rate bound
// all three, which is why the next step is a thread-block-cluster version, not another single-CTA algorithm.
#pragma once
#include "stf_common.cuh"

#define BS_THREADS 1024
#define BS_WARPS 32
#define BS_ITEMS 11                       // elements per thread
#define BS_CAP (BS_THREADS * BS_ITEMS)    // 11264
#define BS_CNT_STRIDE 257

#ifdef STF_BS_DEBUG
__device__ long long g_bs_dbg[64];
#define BS_STAMP(i) do { if (threadIdx.x == 0) g_bs_dbg[i] = clock64(); } while (0)
#else
#define BS_STAMP(i) do { } while (0)
#endif
struct BlockSortSmem {
    uint64_t buf[BS_CAP];
    uint32_t cnt[BS_WARPS * BS_CNT_STRIDE];
    uint32_t wsum[32];
    unsigned long long wor[BS_WARPS];
    uint32_t total;
    int shift[9], npass;
};
#define BS_SMEM_BYTES (sizeof(BlockSortSmem))

__device__ __forceinline__ uint32_t bs_block_exclusive(uint32_t v, uint32_t* wsum, uint32_t* total) {
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
        if (lane == 31 && total) *total = xi;
    }
    __syncthreads();
    return wsum[w] + incl - v;
}

// geometry shared by load / pass / store: warp w owns elements [w*per, (w+1)*per), chunk c = 32 consecutive ones;
// element index of (thread, chunk c) = bs_index(n, c)
__device__ __forceinline__ int bs_chunks(int n) { return ((n + 31) / 32 + BS_WARPS - 1) / BS_WARPS; }
__device__ __forceinline__ int bs_index(int n, int c) { return (int)(threadIdx.x >> 5) * (bs_chunks(n) * 32) + c * 32 + (int)(threadIdx.x & 31); }

// chooses the digit windows: 8-bit windows covering every bit of [bit_lo, bit_hi) in which two elements differ
__device__ __forceinline__ void bs_plan(BlockSortSmem& S, const uint64_t (&x)[BS_ITEMS], int n, int bit_lo, int bit_hi) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int ch = bs_chunks(n), per = ch * 32;
    unsigned long long first = 0, vary = 0;
    bool have = false;
#pragma unroll
    for (int c = 0; c < BS_ITEMS; c++) {
        int e = w * per + c * 32 + lane;
        if (c < ch && e < n) { if (!have) { first = x[c]; have = true; } vary |= x[c] ^ first; }
    }
    // differences against the thread's own first element; threads are tied together through element 0 of the block
    if (threadIdx.x == 0) S.buf[0] = first;
    __syncthreads();
    if (have) vary |= first ^ S.buf[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vary |= __shfl_xor_sync(0xffffffffu, vary, o);
    if (lane == 0) S.wor[w] = vary;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long V = 0;
        for (int i = 0; i < BS_WARPS; i++) V |= S.wor[i];
        if (bit_hi < 64) V &= (1ull << bit_hi) - 1;
        V &= ~((1ull << bit_lo) - 1);
        int np = 0;
        while (V) { int p = __ffsll((long long)V) - 1; S.shift[np++] = p; V &= (p + 8 >= 64) ? 0ull : ~((1ull << (p + 8)) - 1); }
        S.npass = np;
    }
    __syncthreads();
}

// one stable 8-bit pass; on return the elements are in S.buf in sorted order (and, if `reload`, back in x)
__device__ __forceinline__ void bs_pass(BlockSortSmem& S, uint64_t (&x)[BS_ITEMS], int n, int shift, bool reload) {
    const unsigned FULLM = 0xffffffffu;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int ch = bs_chunks(n), per = ch * 32;
    uint32_t* my = S.cnt + w * BS_CNT_STRIDE;
    BS_STAMP(8);
    for (int d = lane; d < 256; d += 32) my[d] = 0;
    __syncwarp();
    uint32_t off[BS_ITEMS];
    unsigned lt = (1u << lane) - 1;
#pragma unroll
    for (int c = 0; c < BS_ITEMS; c++) {
        if (c < ch) { // warp-uniform
            int e = w * per + c * 32 + lane;
            bool valid = e < n;
            uint32_t d = (uint32_t)(x[c] >> shift) & 255u;
            uint32_t peers = __ballot_sync(FULLM, valid); // eight ballots: measured faster than match.any here (many distinct digits)
            if (!valid) peers = ~peers;
#pragma unroll
            for (int b = 0; b < 8; b++) { uint32_t bal = __ballot_sync(FULLM, (d >> b) & 1u); peers &= ((d >> b) & 1u) ? bal : ~bal; }
            int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (valid && lane == leader) { old = my[d]; my[d] = old + __popc(peers); }
            old = __shfl_sync(FULLM, old, leader);
            off[c] = old + __popc(peers & lt);
            __syncwarp();
        }
    }
    __syncthreads();
    BS_STAMP(9);
    { // exclusive scan over (digit, warp): thread t owns digit t>>2, warps (t&3)*8 .. +8
        int d = threadIdx.x >> 2, w0 = (threadIdx.x & 3) * 8;
        uint32_t loc[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { loc[j] = S.cnt[(w0 + j) * BS_CNT_STRIDE + d]; sum += loc[j]; }
        uint32_t run = bs_block_exclusive(sum, S.wsum, nullptr);
#pragma unroll
        for (int j = 0; j < 8; j++) { S.cnt[(w0 + j) * BS_CNT_STRIDE + d] = run; run += loc[j]; }
    }
    __syncthreads();
    BS_STAMP(10);
#pragma unroll
    for (int c = 0; c < BS_ITEMS; c++) {
        if (c < ch) {
            int e = w * per + c * 32 + lane;
            if (e < n) { uint32_t d = (uint32_t)(x[c] >> shift) & 255u; S.buf[my[d] + off[c]] = x[c]; }
        }
    }
    __syncthreads();
    BS_STAMP(11);
    if (reload) {
#pragma unroll
        for (int c = 0; c < BS_ITEMS; c++) {
            int e = w * per + c * 32 + lane;
            if (c < ch && e < n) x[c] = S.buf[e];
        }
        __syncthreads();
    }
    BS_STAMP(12);
}

// sorts the n elements held in x (thread, chunk c <-> element bs_index(n, c)) by bits [bit_lo, bit_hi); the result is
// left in S.buf[0..n) (all threads synchronised)
__device__ __forceinline__ void bs_sort(BlockSortSmem& S, uint64_t (&x)[BS_ITEMS], int n, int bit_lo, int bit_hi) {
    BS_STAMP(0);
    bs_plan(S, x, n, bit_lo, bit_hi);
    int np = S.npass;
    BS_STAMP(1);
#ifdef STF_BS_DEBUG
    if (threadIdx.x == 0) { g_bs_dbg[3] = np; g_bs_dbg[4] = n; }
#endif
    if (np == 0) { // nothing differs: already sorted
#pragma unroll
        for (int c = 0; c < BS_ITEMS; c++) { int e = bs_index(n, c); if (c < bs_chunks(n) && e < n) S.buf[e] = x[c]; }
        __syncthreads();
        return;
    }
    for (int p = 0; p < np; p++) bs_pass(S, x, n, S.shift[p], p + 1 < np);
    BS_STAMP(2);
}
