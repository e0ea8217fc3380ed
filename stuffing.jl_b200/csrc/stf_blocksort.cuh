// stf_blocksort.cuh — latency-oriented sort of one CTA (1024 threads) over <= BS_CAP packed 64-bit words.
//
// Single-scene rounds (10k objects, a few thousand hits) are latency-bound: a multi-launch radix sort costs three
// launches per digit for half a megabyte of data.  Here the whole sort is a device function that fused kernels call.
// One SM does all the work and every block-wide phase costs a few hundred cycles of barriers whatever n is, so the
// design minimises the number of PHASES: a merge sort on the full 64-bit word — thread t sorts its own
// ipt = ceil(n/1024) elements in registers (odd-even network), then ten merge levels double the run length; in every
// level a thread finds its slice of the output with a merge-path binary search and merges ipt elements sequentially
// out of shared memory.  Cost ~ 10 x (search + ipt steps + 2 barriers), independent of how many key bits differ
// (measured: an LSD radix sort with per-thread counters or warp votes needed 6-9 passes of 6-11k cycles each).
// Callers pack (key << tie_bits | tie) so that words are unique: the order is total and deterministic.
#pragma once
#include "stf_common.cuh"

#define BS_THREADS 1024
#define BS_WARPS 32
#define BS_ITEMS 11                       // elements per thread
#define BS_CAP (BS_THREADS * BS_ITEMS)    // 11264
#define BS_PAD 0xFFFFFFFFFFFFFFFFull      // sorts last; no caller produces it

#ifdef STF_BS_DEBUG
__device__ long long g_bs_dbg[64];
#define BS_STAMP(i) do { if (threadIdx.x == 0) g_bs_dbg[i] = clock64(); } while (0)
#else
#define BS_STAMP(i) do { } while (0)
#endif
struct BlockSortSmem {
    uint64_t buf[BS_CAP];
    uint32_t wsum[32];
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

// geometry shared by load / sort / store: thread t owns elements [t*ipt, (t+1)*ipt)
__device__ __forceinline__ int bs_ipt(int n) { return max(1, (n + BS_THREADS - 1) / BS_THREADS); }
__device__ __forceinline__ void bs_ce(uint64_t& a, uint64_t& b) { if (b < a) { uint64_t t = a; a = b; b = t; } }

// sorts the n words held in x (thread t: elements t*ipt .. t*ipt+ipt-1 in x[0..ipt)) ascending; the result is left in
// S.buf[0..n) (all threads synchronised); x is clobbered
__device__ __forceinline__ void bs_sort(BlockSortSmem& S, uint64_t (&x)[BS_ITEMS], int n) {
    BS_STAMP(0);
    const int ipt = bs_ipt(n), t = threadIdx.x, e0 = t * ipt;
#pragma unroll
    for (int i = 0; i < BS_ITEMS; i++) if (i >= ipt || e0 + i >= n) x[i] = BS_PAD;
    // ---- every thread sorts its own elements: odd-even transposition network (padding sinks to the end)
    if (ipt <= 4) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = r & 1; i + 1 < 4; i += 2) bs_ce(x[i], x[i + 1]);
        }
    } else {
#pragma unroll
        for (int r = 0; r < BS_ITEMS; r++) {
#pragma unroll
            for (int i = r & 1; i + 1 < BS_ITEMS; i += 2) bs_ce(x[i], x[i + 1]);
        }
    }
#pragma unroll
    for (int i = 0; i < BS_ITEMS; i++) if (i < ipt) S.buf[e0 + i] = x[i];
    __syncthreads();
    BS_STAMP(1);
    // ---- merge levels: runs of `len` elements (tpr threads each) are merged pairwise
    for (int tpr = 1; tpr < BS_THREADS; tpr <<= 1) {
        const int len = tpr * ipt;
        const int r = t & (2 * tpr - 1), base = (t - r) * ipt;
        const uint64_t* A = S.buf + base; const uint64_t* B = A + len;
        const int diag = r * ipt;
        // merge path: i = number of A elements among the first `diag` outputs (ties take A first: stable)
        int lo = max(0, diag - len), hi = min(diag, len);
        while (lo < hi) { int mid = (lo + hi) >> 1; if (A[mid] <= B[diag - 1 - mid]) lo = mid + 1; else hi = mid; }
        int i = lo, j = diag - lo;
        uint64_t a = i < len ? A[i] : BS_PAD, b = j < len ? B[j] : BS_PAD;
#pragma unroll
        for (int k = 0; k < BS_ITEMS; k++) {
            if (k < ipt) {
                bool ta = j >= len || (i < len && a <= b);
                x[k] = ta ? a : b;
                if (ta) { i++; a = i < len ? A[i] : BS_PAD; } else { j++; b = j < len ? B[j] : BS_PAD; }
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BS_ITEMS; k++) if (k < ipt) S.buf[base + diag + k] = x[k];
        __syncthreads();
    }
    BS_STAMP(2);
#ifdef STF_BS_DEBUG
    if (threadIdx.x == 0) { g_bs_dbg[4] = n; }
#endif
}
