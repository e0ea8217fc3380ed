// stf_sort.cuh — throughput path: multi-launch LSD radix sort (stable, 8-bit digits), order-preserving compaction,
// exclusive scan.  Replaces the Dict / linked-list bucketing of HashSpacialQTree and LinkedSpacialQTree
// (qtrees.jl:411-415, 501-538) and the final sort! of the many-body calls (qtree_functions.jl:252) when a scene or a
// batch of scenes is too large for the single-CTA kernels of stf_finalize.cuh.
//
// * Digits are assembled only from key bits that actually vary (a host-built DigitPass per pass: up to 8 bit runs
//   gathered into one 8-bit digit), so a 49..61-bit cell key costs ceil(varying bits / 8) passes, not 7.
// * Granularity: one WARP owns a contiguous slice of the input and a private 256-bin histogram in shared memory.
//   Ranks inside a 32-key chunk come from __match_any_sync, so equal digits keep their input order (stable).
// * Three launches per pass: histogram, scan (one CTA per DIGIT scans that digit's per-warp counts; the 256 digit
//   totals are scanned again by every scatter CTA), scatter.
#pragma once
#include <vector>
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

// one 8-bit digit = up to 8 runs of key bits: digit = sum ((key >> shift) & mask) << out
struct DigitPass { int nseg; int shift[8]; uint32_t mask[8]; int out[8]; };
__host__ __device__ __forceinline__ uint32_t digit_of(uint64_t key, const DigitPass& P) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) if (i < P.nseg) d |= ((uint32_t)(key >> P.shift[i]) & P.mask[i]) << P.out[i];
    return d;
}
// passes that sort by exactly the bits set in `bits` (LSB-first): the order equals the order by the full key whenever
// all other bits are equal across the keys
static inline std::vector<DigitPass> plan_digits(uint64_t bits) {
    std::vector<DigitPass> passes;
    DigitPass cur = {}; int fill = 0;
    int b = 0;
    while (b < 64) {
        if (!((bits >> b) & 1)) { b++; continue; }
        int e = b; while (e < 64 && ((bits >> e) & 1)) e++; // run [b, e)
        while (b < e) {
            int take = std::min(e - b, 8 - fill);
            cur.shift[cur.nseg] = b; cur.mask[cur.nseg] = (1u << take) - 1; cur.out[cur.nseg] = fill; cur.nseg++;
            fill += take; b += take;
            if (fill == 8) { passes.push_back(cur); cur = DigitPass(); fill = 0; }
        }
    }
    if (fill) passes.push_back(cur);
    return passes;
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const uint64_t* __restrict__ keys, int64_t n, DigitPass P, int per_warp, int nvw, uint32_t* __restrict__ hist) {
    pdl_sync();
    __shared__ uint32_t cnt[SORT_WARPS_PER_BLOCK][256];
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int vw = blockIdx.x * SORT_WARPS_PER_BLOCK + w;
    for (int d = lane; d < 256; d += 32) cnt[w][d] = 0;
    __syncwarp();
    int64_t beg = (int64_t)vw * per_warp, end = min(n, beg + per_warp);
    for (int64_t i0 = beg + lane; i0 < end; i0 += 32 * 8) { // eight loads in flight per lane
        uint64_t k[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { int64_t i = i0 + 32 * j; k[j] = i < end ? keys[i] : 0ull; }
#pragma unroll
        for (int j = 0; j < 8; j++) if (i0 + 32 * j < end) atomicAdd(&cnt[w][digit_of(k[j], P)], 1u);
    }
    __syncwarp();
    for (int d = lane; d < 256; d += 32) hist[(size_t)d * nvw + vw] = cnt[w][d];
}
// CTA d: exclusive scan of digit d's per-warp counts (in place), digit total -> dtot[d]
__global__ void __launch_bounds__(256) k_sort_scan_digit(uint32_t* __restrict__ hist, int nvw, uint32_t* __restrict__ dtot) {
    pdl_sync();
    __shared__ uint32_t wsum[8];
    uint32_t* row = hist + (size_t)blockIdx.x * nvw;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int per = (nvw + 255) / 256, beg = threadIdx.x * per, end = min(nvw, beg + per);
    uint32_t s = 0;
    for (int i = beg; i < end; i++) s += row[i];
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    uint32_t off = 0, tot = 0;
    for (int i = 0; i < 8; i++) { if (i < w) off += wsum[i]; tot += wsum[i]; }
    uint32_t run = off + incl - s;
    for (int i = beg; i < end; i++) { uint32_t v = row[i]; row[i] = run; run += v; }
    if (threadIdx.x == 0) dtot[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n, DigitPass P,
                                                               int per_warp, int nvw, const uint32_t* __restrict__ hist, const uint32_t* __restrict__ dtot,
                                                               uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    pdl_sync();
    __shared__ uint32_t base[SORT_WARPS_PER_BLOCK][256];
    __shared__ uint32_t dbase[256];
    __shared__ uint32_t wsum[8];
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    { // exclusive scan of the 256 digit totals (SORT_THREADS == 256: one digit per thread)
        uint32_t v = dtot[threadIdx.x], incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        uint32_t off = 0;
        for (int i = 0; i < w; i++) off += wsum[i];
        dbase[threadIdx.x] = off + incl - v;
        __syncthreads();
    }
    int vw = blockIdx.x * SORT_WARPS_PER_BLOCK + w;
    for (int d = lane; d < 256; d += 32) base[w][d] = hist[(size_t)d * nvw + vw] + dbase[d];
    __syncwarp();
    int64_t beg = (int64_t)vw * per_warp, end = min(n, beg + per_warp);
    // SC_K chunks of 32 keys per round: the loads and the match.any votes of a round are independent (in flight
    // together); only the running digit bases in shared memory chain one chunk to the next
#define SC_K 8
    const uint32_t lt = (1u << lane) - 1;
    for (int64_t i0 = beg; i0 < end; i0 += 32 * SC_K) {
        uint64_t k[SC_K]; uint32_t v[SC_K], d[SC_K], peers[SC_K], dst[SC_K];
#pragma unroll
        for (int j = 0; j < SC_K; j++) {
            int64_t i = i0 + 32 * j + lane;
            bool valid = i < end;
            k[j] = valid ? keys[i] : 0ull; v[j] = valid ? vals[i] : 0u;
        }
#pragma unroll
        for (int j = 0; j < SC_K; j++) {
            bool valid = i0 + 32 * j + lane < end;
            d[j] = valid ? digit_of(k[j], P) : 256u + lane; // invalid lanes never match
            peers[j] = __match_any_sync(0xffffffffu, d[j]);
        }
#pragma unroll
        for (int j = 0; j < SC_K; j++) {
            bool valid = i0 + 32 * j + lane < end;
            uint32_t rank = __popc(peers[j] & lt);
            dst[j] = valid ? base[w][d[j]] + rank : 0u;
            __syncwarp();
            if (valid && rank == 0) base[w][d[j]] += __popc(peers[j]);
            __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < SC_K; j++)
            if (i0 + 32 * j + lane < end) { keys_out[dst[j]] = k[j]; vals_out[dst[j]] = v[j]; }
    }
#undef SC_K
}

// ---- one launch per pass ("onesweep"): the three launches of a pass above (per-warp histograms, per-digit scan, scatter)
// cost ~8 us each on mid-size inputs whatever the work, and a batch step runs 17 passes.  Here ONE upfront kernel counts
// the digits of EVERY pass (the multiset of keys does not change between passes) and clears the tile states; a pass is then
// one kernel: a CTA takes the next tile of 2048 keys (atomic ticket, so a tile only waits for tiles that already run),
// ranks its keys stably (match.any per 32-key chunk + warp-private running counts + warp prefixes), publishes its 256
// digit counts, obtains the counts of all earlier tiles by DECOUPLED LOOK-BACK (thread d walks back over the states of
// digit d until it meets an inclusive prefix) and scatters.  state = pass(8) | kind(2: 1 = tile count, 2 = inclusive prefix) |
// value(54).  Same stable order as the three-launch version, hence bit-identical results.
#define OS_THREADS 256
#define OS_KPT 8
#define OS_TILE (OS_THREADS * OS_KPT)
#define OS_MAX_PASSES 8
struct OsPasses { int np; DigitPass P[OS_MAX_PASSES]; };
__global__ void __launch_bounds__(256) k_os_hist(const uint64_t* __restrict__ keys, int64_t n, OsPasses Q, uint32_t* __restrict__ ghist /* np x 256 */,
                                                 unsigned long long* __restrict__ status, int64_t nstatus, unsigned int* __restrict__ tickets /* OS_MAX_PASSES */) {
    pdl_sync();
    __shared__ uint32_t cnt[OS_MAX_PASSES][256];
    for (int i = threadIdx.x; i < OS_MAX_PASSES * 256; i += 256) (&cnt[0][0])[i] = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nstatus; i += (int64_t)gridDim.x * 256) status[i] = 0ull;
    if (blockIdx.x == 0 && threadIdx.x < OS_MAX_PASSES) tickets[threadIdx.x] = 0u;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const uint64_t k = keys[i];
        for (int p = 0; p < Q.np; p++) atomicAdd(&cnt[p][digit_of(k, Q.P[p])], 1u);
    }
    __syncthreads();
    for (int p = 0; p < Q.np; p++) { uint32_t c = cnt[p][threadIdx.x]; if (c) atomicAdd(ghist + p * 256 + threadIdx.x, c); }
}
__global__ void __launch_bounds__(OS_THREADS) k_os_pass(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n, DigitPass P, int pass,
                                                        const uint32_t* __restrict__ ghist, unsigned long long* __restrict__ status, unsigned int* __restrict__ ticket,
                                                        uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    pdl_sync();
    __shared__ uint32_t wcnt[OS_THREADS / 32][256]; // warp-private running counts, then exclusive warp prefixes
    __shared__ uint32_t dbase[256];
    __shared__ uint32_t wsum[8];
    __shared__ unsigned int s_tile;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long TAG = (unsigned long long)(pass + 1) << 56;
    // exclusive scan of the pass's 256 global digit totals (one digit per thread)
    uint32_t gbase;
    {
        uint32_t v = ghist[pass * 256 + threadIdx.x], incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        uint32_t off = 0;
        for (int i = 0; i < w; i++) off += wsum[i];
        gbase = off + incl - v;
    }
    const int64_t ntiles = (n + OS_TILE - 1) / OS_TILE;
    const uint32_t lt = (1u << lane) - 1;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        for (int d = lane; d < 256; d += 32) wcnt[w][d] = 0;
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= ntiles) return;
        const int64_t beg = tile * OS_TILE + (int64_t)w * (32 * OS_KPT), end = min(n, tile * OS_TILE + OS_TILE);
        uint64_t k[OS_KPT]; uint32_t v[OS_KPT], d[OS_KPT], peers[OS_KPT], loc[OS_KPT];
#pragma unroll
        for (int j = 0; j < OS_KPT; j++) {
            const int64_t i = beg + 32 * j + lane;
            const bool valid = i < end;
            k[j] = valid ? keys[i] : 0ull; v[j] = valid ? vals[i] : 0u;
        }
#pragma unroll
        for (int j = 0; j < OS_KPT; j++) {
            const bool valid = beg + 32 * j + lane < end;
            d[j] = valid ? digit_of(k[j], P) : 256u + lane; // invalid lanes never match
            peers[j] = __match_any_sync(0xffffffffu, d[j]);
        }
#pragma unroll
        for (int j = 0; j < OS_KPT; j++) { // rank inside the warp's 256 keys: running count of the digit + rank in the chunk
            const bool valid = beg + 32 * j + lane < end;
            const uint32_t rank = __popc(peers[j] & lt);
            loc[j] = valid ? wcnt[w][d[j]] + rank : 0u;
            __syncwarp();
            if (valid && rank == 0) wcnt[w][d[j]] += __popc(peers[j]);
            __syncwarp();
        }
        __syncthreads();
        { // digit threadIdx.x: warp counts -> exclusive warp prefixes; tile count published; look-back over the earlier tiles
            const int dg = threadIdx.x;
            uint32_t run = 0;
#pragma unroll
            for (int i = 0; i < OS_THREADS / 32; i++) { uint32_t c = wcnt[i][dg]; wcnt[i][dg] = run; run += c; }
            unsigned long long* my = status + (size_t)tile * 256 + dg;
            unsigned long long excl = 0;
            if (tile == 0) {
                asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(my), "l"(TAG | (2ull << 54) | (unsigned long long)run) : "memory");
            } else {
                asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(my), "l"(TAG | (1ull << 54) | (unsigned long long)run) : "memory");
                for (int64_t t2 = tile - 1; t2 >= 0; t2--) {
                    const unsigned long long* st = status + (size_t)t2 * 256 + dg;
                    unsigned long long sv;
                    do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(sv) : "l"(st) : "memory"); } while ((sv >> 56) != (unsigned long long)(pass + 1) || ((sv >> 54) & 3ull) == 0ull);
                    excl += sv & ((1ull << 54) - 1ull);
                    if (((sv >> 54) & 3ull) == 2ull) break;
                }
                asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(my), "l"(TAG | (2ull << 54) | (excl + (unsigned long long)run)) : "memory");
            }
            dbase[dg] = gbase + (uint32_t)excl;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < OS_KPT; j++)
            if (beg + 32 * j + lane < end) { const uint32_t dst = dbase[d[j]] + wcnt[w][d[j]] + loc[j]; keys_out[dst] = k[j]; vals_out[dst] = v[j]; }
    }
}

// Sorts (keys, vals) by the key bits set in `bits`.  Buffers ping-pong; returns 0 if the result is in (keys, vals),
// 1 if it is in (keys_alt, vals_alt).  `hist` holds >= radix_sort_hist_words(n) uint32 of scratch.
template <typename... KArgs, typename... Args>
static inline void sort_launch(bool pdl, cudaStream_t st, void (*kern)(KArgs...), int grid, int block, Args... args) {
    if (!pdl) { kern<<<grid, block, 0, st>>>(args...); return; }
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = st;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, args...);
}
static inline size_t os_status_words(int64_t n) { return (size_t)((n + OS_TILE - 1) / OS_TILE) * 256; } // u64 each
static inline int radix_sort_pairs(cudaStream_t st, int64_t* launches, uint64_t* keys, uint32_t* vals, uint64_t* keys_alt, uint32_t* vals_alt,
                                   int64_t n, uint64_t bits, uint32_t* hist, bool pdl = false, bool* after_kernel = nullptr, bool three_launch = false) {
    // *after_kernel: the stream's last operation is a kernel (a copy or memset before a programmatic launch is not waited for)
    bool pdl_first = after_kernel && *after_kernel;
    if (n <= 1) return 0;
    std::vector<DigitPass> passes = plan_digits(bits);
    int flip = 0;
    // one launch per pass for mid-size inputs (measured on B200: C4 sort stage 0.087 -> 0.079 ms, a 512-scene batch 0.305 -> 0.264 ms
    // over its three sorts; at 0.5 M keys the look-back over 260 concurrent tiles costs what the saved launches gain, and
    // 8192-key tiles on 1024-thread CTAs were slower at every size)
    if (!three_launch && (int)passes.size() <= OS_MAX_PASSES && n <= (1 << 18)) {
        if (passes.empty()) return 0;
        // scratch layout (uint32 words): [tile states: 2 words each][global digit counts: np x 256][tickets: OS_MAX_PASSES]
        const size_t nst = os_status_words(n);
        unsigned long long* status = reinterpret_cast<unsigned long long*>(hist);
        uint32_t* ghist = hist + 2 * nst;
        unsigned int* tickets = reinterpret_cast<unsigned int*>(ghist + 256 * OS_MAX_PASSES);
        OsPasses Q; Q.np = (int)passes.size();
        for (int p = 0; p < Q.np; p++) Q.P[p] = passes[p];
        cudaMemsetAsync(ghist, 0, sizeof(uint32_t) * 256 * OS_MAX_PASSES, st); // (a memset, not a kernel: the next launch is a plain one)
        const int hb = (int)std::min<int64_t>(std::max<int64_t>(1, (n + 2047) / 2048), 148 * 4);
        sort_launch(false, st, k_os_hist, hb, 256, (const uint64_t*)keys, n, Q, ghist, status, (int64_t)nst, tickets);
        const int grid = (int)std::min<int64_t>((n + OS_TILE - 1) / OS_TILE, 148 * 3);
        for (int p = 0; p < Q.np; p++) {
            uint64_t* kin = flip ? keys_alt : keys; uint32_t* vin = flip ? vals_alt : vals;
            uint64_t* kout = flip ? keys : keys_alt; uint32_t* vout = flip ? vals : vals_alt;
            sort_launch(pdl, st, k_os_pass, grid, OS_THREADS, (const uint64_t*)kin, (const uint32_t*)vin, n, passes[p], p, (const uint32_t*)ghist, status, tickets + p, kout, vout);
            flip ^= 1;
        }
        if (launches) *launches += 1 + Q.np;
        if (after_kernel) *after_kernel = true;
        return flip;
    }
    SortPlan p = sort_plan(n);
    uint32_t* dtot = hist + (size_t)256 * p.nvw;
    for (const DigitPass& P : passes) {
        uint64_t* kin = flip ? keys_alt : keys; uint32_t* vin = flip ? vals_alt : vals;
        uint64_t* kout = flip ? keys : keys_alt; uint32_t* vout = flip ? vals : vals_alt;
        sort_launch(pdl && pdl_first, st, k_sort_hist, p.blocks, SORT_THREADS, (const uint64_t*)kin, n, P, p.per_warp, p.nvw, hist);
        sort_launch(pdl, st, k_sort_scan_digit, 256, 256, hist, p.nvw, dtot);
        sort_launch(pdl, st, k_sort_scatter, p.blocks, SORT_THREADS, (const uint64_t*)kin, (const uint32_t*)vin, n, P, p.per_warp, p.nvw, (const uint32_t*)hist, (const uint32_t*)dtot, kout, vout);
        flip ^= 1; if (launches) *launches += 3;
        pdl_first = true; if (after_kernel) *after_kernel = true;
    }
    return flip;
}
static inline uint64_t bit_range(int lo, int hi) { return (hi >= 64 ? ~0ull : ((1ull << hi) - 1)) & ~((1ull << lo) - 1); }
static inline size_t radix_sort_hist_words(int64_t n) { // scratch of either variant
    return std::max((size_t)256 * sort_plan(n).nvw + 256, 2 * os_status_words(n) + 256 * OS_MAX_PASSES + OS_MAX_PASSES + 16);
}

// exclusive scan of `m` uint32 by ONE block (small m only: per-block counts of a compaction)
__global__ void __launch_bounds__(1024) k_scan_u32(uint32_t* __restrict__ data, int64_t m, uint32_t* __restrict__ total) {
    pdl_sync();
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

// ---- order-preserving compaction of the 4-slot record table (INVALID keys dropped) + the OR of all key differences
#define CMP_BLOCK 1024
__global__ void __launch_bounds__(CMP_BLOCK) k_compact_count(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ blockcnt) {
    pdl_sync();
    int64_t i = (int64_t)blockIdx.x * CMP_BLOCK + threadIdx.x;
    int c = __syncthreads_count(i < n && keys[i] != STF_KEY_INVALID);
    if (threadIdx.x == 0) blockcnt[blockIdx.x] = (uint32_t)c;
}
__global__ void __launch_bounds__(CMP_BLOCK) k_compact_write(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n, const uint32_t* __restrict__ blockoff,
                                                             uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    pdl_sync();
    __shared__ uint32_t wsum[32];
    int64_t i = (int64_t)blockIdx.x * CMP_BLOCK + threadIdx.x;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t k = i < n ? keys[i] : STF_KEY_INVALID;
    bool v = k != STF_KEY_INVALID;
    unsigned m = __ballot_sync(0xffffffffu, v);
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    uint32_t off = blockoff[blockIdx.x];
    for (int j = 0; j < w; j++) off += wsum[j];
    if (v) { uint32_t dst = off + __popc(m & ((1u << lane) - 1)); keys_out[dst] = k; vals_out[dst] = vals[i]; }
}
// vary |= OR over i < n of keys[i] ^ keys[0]  (n on the device): the bits the sort has to look at
__global__ void __launch_bounds__(256) k_vary_or(const uint64_t* __restrict__ keys, const unsigned long long* __restrict__ n_dev, int64_t cap, unsigned long long* __restrict__ vary) {
    pdl_sync();
    __shared__ unsigned long long wor[8];
    int64_t n = min((int64_t)*n_dev, cap);
    unsigned long long x = 0;
    if (n > 0) {
        uint64_t ref = keys[0];
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x |= keys[i] ^ ref;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x |= __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) wor[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long t = 0; for (int j = 0; j < 8; j++) t |= wor[j]; if (t) atomicOr(vary, t); }
}
