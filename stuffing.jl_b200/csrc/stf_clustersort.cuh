// stf_clustersort.cuh — the latency-path sort on a THREAD-BLOCK CLUSTER: 8 CTAs (8 SMs of one GPC) sort up to 16 384
// packed 64-bit words together, exchanging histograms and elements through distributed shared memory.
//
// Why: measured on B200, every single-CTA sort of 3k-11k elements (ballot radix, per-thread-counter radix, merge path)
// needs 20-100k cycles because ONE SM's issue rate and shared-memory bandwidth bound it (stf_blocksort.cuh).  A cluster
// gives 8x the issue slots and shared-memory bandwidth, and cluster.sync (~380 cycles) is cheap enough to pay twice per
// pass.  Algorithm = the same stable LSD radix sort over the bit windows that vary:
//   CTA c owns the sorted positions [c*P, (c+1)*P); elements live in registers between passes;
//   per pass: ballot ranks + warp-private histograms (local) -> per-digit totals published to every CTA through DSMEM
//   -> cluster.sync -> every CTA derives its digit bases -> elements are scattered straight into the owner CTA's shared
//   memory (remote 8-byte stores) -> cluster.sync -> reload.
#pragma once
#include <cooperative_groups.h>
#include "stf_common.cuh"
namespace cg = cooperative_groups;

#define CS_CTAS 8
#define CS_THREADS 1024
#define CS_WARPS 32
#define CS_ITEMS 2                               // elements per thread
#define CS_SLICE (CS_THREADS * CS_ITEMS)         // 2048 elements per CTA
#define CS_CAP (CS_CTAS * CS_SLICE)              // 16384
#define CS_STRIDE 257

struct ClusterSortSmem {
    uint64_t buf[CS_SLICE];
    uint32_t cnt[CS_WARPS * CS_STRIDE];          // warp-private histograms, then within-CTA warp prefixes
    uint32_t gt[CS_CTAS][256];                   // digit totals of every CTA of the cluster (filled remotely)
    uint32_t dbase[256];
    uint32_t wsum[32];
    unsigned long long pf[CS_CTAS], pv[CS_CTAS]; // first element / OR of differences of every CTA (filled remotely)
    int shift[9], npass;
};
#define CS_SMEM_BYTES (sizeof(ClusterSortSmem))

// slice geometry: P elements per CTA (multiple of 32), CTA c holds nl = clamp(n - c*P, 0, P) of them;
// (thread, item j) <-> local element w*per + j*32 + lane with per = CS_ITEMS*32
__device__ __forceinline__ int cs_slice(int n) { int p = (n + CS_CTAS - 1) / CS_CTAS; return max(32, (p + 31) & ~31); }
__device__ __forceinline__ int cs_local(int j) { return (int)(threadIdx.x >> 5) * (CS_ITEMS * 32) + j * 32 + (int)(threadIdx.x & 31); }

__device__ __forceinline__ uint32_t cs_block_exclusive(uint32_t v, uint32_t* wsum) {
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
    }
    __syncthreads();
    return wsum[w] + incl - v;
}

// sorts the n words distributed over the cluster (CTA c: x[j] = element c*P + cs_local(j), valid while cs_local(j) < nl)
// by bits [bit_lo, bit_hi); on return CTA c holds the sorted positions [c*P, c*P + nl) in S.buf[0..nl) (cluster-synchronised)
__device__ __forceinline__ void cs_sort(ClusterSortSmem& S, uint64_t (&x)[CS_ITEMS], int n, int bit_lo, int bit_hi) {
    const unsigned FULLM = 0xffffffffu;
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const int P = cs_slice(n), nl = min(max(n - c * P, 0), P);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // ---- digit windows: OR of all differences, cluster-wide
    {
        unsigned long long first = 0, vary = 0; bool have = false;
#pragma unroll
        for (int j = 0; j < CS_ITEMS; j++) if (cs_local(j) < nl) { if (!have) { first = x[j]; have = true; } vary |= x[j] ^ first; }
        if (threadIdx.x == 0) S.buf[0] = first; // thread 0 holds local element 0 (valid iff nl > 0)
        __syncthreads();
        if (have) vary |= first ^ S.buf[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vary |= __shfl_xor_sync(FULLM, vary, o);
        unsigned long long* wor = reinterpret_cast<unsigned long long*>(S.cnt);
        if (lane == 0) wor[w] = vary;
        __syncthreads();
        if (threadIdx.x < CS_CTAS) { // publish (first, vary) of this CTA to CTA `threadIdx.x`
            unsigned long long V = 0;
            for (int i = 0; i < CS_WARPS; i++) V |= wor[i];
            ClusterSortSmem* R = cluster.map_shared_rank(&S, threadIdx.x);
            R->pf[c] = S.buf[0]; R->pv[c] = nl > 0 ? (V | 0x8000000000000000ull) : 0ull; // top bit = "this CTA has elements"
        }
        cluster.sync();
        if (threadIdx.x == 0) {
            unsigned long long V = 0, ref = 0; bool got = false;
            for (int r = 0; r < CS_CTAS; r++) if (S.pv[r]) { if (!got) { ref = S.pf[r]; got = true; } V |= (S.pv[r] & 0x7FFFFFFFFFFFFFFFull) | (S.pf[r] ^ ref); }
            if (bit_hi < 64) V &= (1ull << bit_hi) - 1;
            V &= ~((1ull << bit_lo) - 1);
            int np = 0;
            while (V) { int p = __ffsll((long long)V) - 1; S.shift[np++] = p; V &= (p + 8 >= 64) ? 0ull : ~((1ull << (p + 8)) - 1); }
            S.npass = np;
        }
        __syncthreads();
    }
    const int np = S.npass;
    for (int pass = 0; pass < np; pass++) {
        const int shift = S.shift[pass];
        uint32_t* my = S.cnt + w * CS_STRIDE;
        for (int d = lane; d < 256; d += 32) my[d] = 0;
        __syncwarp();
        uint32_t off[CS_ITEMS];
        const unsigned lt = (1u << lane) - 1;
#pragma unroll
        for (int j = 0; j < CS_ITEMS; j++) {
            bool valid = cs_local(j) < nl;
            uint32_t d = (uint32_t)(x[j] >> shift) & 255u;
            uint32_t peers = __ballot_sync(FULLM, valid);
            if (!valid) peers = ~peers;
#pragma unroll
            for (int b = 0; b < 8; b++) { uint32_t bal = __ballot_sync(FULLM, (d >> b) & 1u); peers &= ((d >> b) & 1u) ? bal : ~bal; }
            int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (valid && lane == leader) { old = my[d]; my[d] = old + __popc(peers); }
            old = __shfl_sync(FULLM, old, leader);
            off[j] = old + __popc(peers & lt);
            __syncwarp();
        }
        __syncthreads();
        if (threadIdx.x < 256) { // digit d: warp prefixes inside this CTA, the CTA total to every CTA of the cluster
            const int d = threadIdx.x;
            uint32_t run = 0;
#pragma unroll 8
            for (int i = 0; i < CS_WARPS; i++) { uint32_t v = S.cnt[i * CS_STRIDE + d]; S.cnt[i * CS_STRIDE + d] = run; run += v; }
#pragma unroll
            for (int r = 0; r < CS_CTAS; r++) cluster.map_shared_rank(&S, r)->gt[c][d] = run;
        }
        cluster.sync();
        { // digit bases: all smaller digits of the whole cluster + the same digit in the CTAs before this one
            uint32_t col = 0, pre = 0;
            if (threadIdx.x < 256) {
#pragma unroll
                for (int r = 0; r < CS_CTAS; r++) { uint32_t v = S.gt[r][threadIdx.x]; col += v; if (r < c) pre += v; }
            }
            uint32_t ex = cs_block_exclusive(col, S.wsum);
            if (threadIdx.x < 256) S.dbase[threadIdx.x] = ex + pre;
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < CS_ITEMS; j++) {
            if (cs_local(j) < nl) {
                uint32_t d = (uint32_t)(x[j] >> shift) & 255u;
                uint32_t dst = S.dbase[d] + my[d] + off[j];
                uint32_t owner = dst / (uint32_t)P;
                cluster.map_shared_rank(&S, owner)->buf[dst - owner * (uint32_t)P] = x[j];
            }
        }
        cluster.sync();
#pragma unroll
        for (int j = 0; j < CS_ITEMS; j++) if (cs_local(j) < nl) x[j] = S.buf[cs_local(j)];
        // (the next pass overwrites S.buf only after its own first cluster.sync, i.e. after every CTA has reloaded)
    }
    if (np == 0) { // nothing differs: the input order is the sorted order
#pragma unroll
        for (int j = 0; j < CS_ITEMS; j++) if (cs_local(j) < nl) S.buf[cs_local(j)] = x[j];
        cluster.sync();
    }
}
