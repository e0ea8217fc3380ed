// stf_pyramid.cuh — subsystem (1): pack + pyramid build / rebuild kernels.
// Replaces PaddedMat/ShiftedQTree construction and buildqtree! (qtrees.jl:117-128,176-237) and the
// rebuild half of shift!/setshift! (qtrees.jl:238-287).
#pragma once
#include "stf_common.cuh"

// ---- level-1 pack: u8 codes (column-major payload, pitch kh) -> 2-bit words -------------------
// word w of a column: 16 payload bytes at any alignment are fetched as 5 aligned words + funnel shifts (the buffer has
// slack); the tail word of a column takes `default` in the fields past kh.  Validates the codes (1..3).
__device__ __forceinline__ uint32_t pack16(const uint8_t* __restrict__ src, uint32_t fm, uint32_t dw, uint32_t& hi_acc, uint32_t& zero_acc) {
    // 16 u8 codes at any byte alignment -> one 2-bit word.  One branch-free path for full and tail words (a warp mixes
    // both): always fetch 16 bytes as 5 aligned words + funnel shifts — a tail word over-reads into the next column / next
    // object / the zeroed slack, all of which have clear high bits — and replace the fields outside `fm` by `default`.
    // 4 bytes -> 8 bits with one multiply: (x & 0x03030303) * (2^24+2^18+2^12+2^6) moves byte k's code to bits 24+2k
    // (the cross terms land below bit 24 or overflow, none carries); three byte-permutes gather the four top bytes.
    const uint32_t* ap = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
    int sh = (int)(reinterpret_cast<uintptr_t>(src) & 3) * 8;
    uint32_t t[5];
#pragma unroll
    for (int k = 0; k < 5; k++) t[k] = __ldg(ap + k);
    uint32_t p[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t x = __funnelshift_r(t[k], t[k + 1], sh);
        hi_acc |= x;
        p[k] = (x & 0x03030303u) * 0x01041040u;
    }
    uint32_t out = __byte_perm(__byte_perm(p[0], p[1], 0x0073), __byte_perm(p[2], p[3], 0x0073), 0x5410);
    zero_acc |= ~(out | (out >> 1)) & 0x55555555u & fm; // codes are 1..3: a zero field among the valid ones is an error
    return (out & fm) | (dw & ~fm);
}
// field mask of word w of a column of kh nodes (all 16 fields, or the first kh - 16w of the last word)
__device__ __forceinline__ uint32_t tail_mask(int kh) { int nv = kh - 16 * (wpc_of(kh) - 1); return nv >= 16 ? 0xFFFFFFFFu : ((1u << (2 * nv)) - 1u); }
__device__ __forceinline__ uint32_t pack_word(const uint8_t* __restrict__ col, int kh, int w, uint32_t d, int& bad) {
    uint32_t hi = 0, z = 0;
    uint32_t r = pack16(col + 16 * w, w == wpc_of(kh) - 1 ? tail_mask(kh) : 0xFFFFFFFFu, dflt_word(d), hi, z);
    bad |= ((hi & 0xFCFCFCFCu) != 0) | (z != 0);
    return r;
}
// one thread per output word of the listed (big) objects; `woff1[i]` = first level-1 word of list entry i in a dense
// numbering (prefix sum), `poff[o]` = byte offset of object o's payload.
__global__ void k_pack_level1(const uint8_t* __restrict__ payload, const uint64_t* __restrict__ poff, const int32_t* __restrict__ list,
                              const uint64_t* __restrict__ woff1, int nlist, const LevelMeta* __restrict__ lm, int L,
                              uint32_t* __restrict__ words, uint64_t total_words, int* __restrict__ err) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_words) return;
    int lo = 0, hi = nlist; // largest entry with woff1[entry] <= i
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (woff1[mid] <= i) lo = mid; else hi = mid; }
    int o = list[lo];
    LevelMeta m = lm[(size_t)o * L];
    int kh = lm_kh(m), wpc = wpc_of(kh);
    uint32_t local = (uint32_t)(i - woff1[lo]);
    int c = local / wpc, w = local - c * wpc;
    int bad = 0;
    uint32_t out = pack_word(payload + poff[o] + (size_t)c * kh, kh, w, lm_dflt(m), bad);
    if (bad) atomicOr(err, 1);
    words[m.off + local] = out;
}

// ---- fused pack + build, small objects (level 1 <= RB_HALF words).  Instruction issue, not bandwidth, limits this
// kernel (a ~1 KB object has as many levels as a mask), so the work is split by shape:
//   phase A, one WARP per object: read the payload once (contiguous, all loads in flight together), pack level 1 into
//            global AND shared memory, build the levels that still have real width out of shared memory / registers;
//   phase B, one THREAD per object (32 objects of the warp at once): from the first level with kh <= 16 and kw <= 8 up
//            to the root the whole level lives in 8 registers; a level is 5 x (two selects per child column, pair-OR,
//            2-bit compress) — ~1 warp instruction per object and level instead of ~100.
// HBM traffic = the payload once + the packed pyramid once.
__device__ __forceinline__ uint32_t sel3(int e, uint32_t a, uint32_t b, uint32_t c) { return e > 0 ? a : (e == 0 ? b : c); }
__device__ __forceinline__ int tail_level(int kh, int kw, int L) { // first level that fits 8 registers (or L)
    int t = 1;
    while ((kh > 16 || kw > 8) && t < L) { kh = kh / 2 + 1; kw = kw / 2 + 1; t++; }
    return t;
}
// Thread-per-object tail of a (re)build: starting from level `from` (whose raster and every raster up to the first
// level T with kh <= 16 and kw <= 8 are valid in memory) one THREAD builds the levels T+1..L of its object: the whole
// level lives in 8 registers, a level is 5 x (two selects per child column, pair-OR, 2-bit compress).
// KEEP: the per-level shifts are state and are read from the metas (lazy batchsteps); otherwise shift_l = shift_{l-1} / 2.
template <bool KEEP> __device__ __forceinline__ void thread_tail_build(uint32_t* __restrict__ words, const LevelMeta* __restrict__ olm, int L, int from) {
    LevelMeta m = olm[from - 1];
    int kh = lm_kh(m), kw = lm_kw(m), rs = m.rs, cs = m.cs, t = from;
    uint32_t off = m.off; const uint32_t d = lm_dflt(m), dw = dflt_word(d);
    while ((kh > 16 || kw > 8) && t < L) {
        off += (uint32_t)(wpc_of(kh) * kw); kh = kh / 2 + 1; kw = kw / 2 + 1; t++;
        if (KEEP) { LevelMeta x = olm[t - 1]; rs = x.rs; cs = x.cs; } else { rs /= 2; cs /= 2; }
    }
    if (!(t < L && kh <= 16 && kw <= 8)) return;
    uint32_t cur[8];
#pragma unroll
    for (int i = 0; i < 8; i++) cur[i] = i < kw ? words[off + i] : dw;
    for (int l = t + 1; l <= L; l++) {
        int prs = rs / 2, pcs = cs / 2; // qtrees.jl:225-226 (÷ truncates toward zero)
        if (KEEP) { LevelMeta x = olm[l - 1]; prs = x.rs; pcs = x.cs; }
        int er = rs - 2 * prs, ec = cs - 2 * pcs, kwp = kw / 2 + 1;
        uint32_t poff_ = off + (uint32_t)kw; // one word per column at these sizes
        uint32_t nxt[5];
#pragma unroll
        for (int pc = 0; pc < 5; pc++) {
#define STF_COL(i) (((i) >= 0 && (i) < 8 && (i) < kw) ? cur[((i) >= 0 && (i) < 8) ? (i) : 0] : dw)
            uint32_t w0 = sel3(ec, STF_COL(2 * pc - 1), STF_COL(2 * pc), STF_COL(2 * pc + 1));
            uint32_t w1 = sel3(ec, STF_COL(2 * pc), STF_COL(2 * pc + 1), STF_COL(2 * pc + 2));
#undef STF_COL
            uint64_t y = ((uint64_t)dw << 32) | (w0 | w1);
            if (er > 0) y = (y << 2) | d; else if (er < 0) y = (y >> 2) | ((uint64_t)d << 62);
            nxt[pc] = compress_pairs(y);
            if (pc < kwp) words[poff_ + pc] = nxt[pc];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) cur[i] = i < 5 ? nxt[i < 5 ? i : 0] : dw;
        off = poff_; kh = kh / 2 + 1; kw = kwp; rs = prs; cs = pcs;
    }
}
// first level >= from that the thread tail can take over (or L)
__device__ __forceinline__ int tail_level_from(int kh, int kw, int L, int from) { // (kh, kw) = dims of level `from`
    int t = from;
    while ((kh > 16 || kw > 8) && t < L) { kh = kh / 2 + 1; kw = kw / 2 + 1; t++; }
    return t;
}
// (measured on B200, 4096 C5 scenes: unroll 1 / 6 CTAs per SM 515 us, 2 / 5 502 us, 4 / 4 520 us, 8 / 3 581 us — the kernel is
// bound by its instruction count (ncu: 642 warp instructions per object, 0.66 issued per scheduler cycle), not by loads in flight)
#ifndef STF_PACK_UNROLL
#define STF_PACK_UNROLL 2
#endif
#ifndef STF_PACK_MINB
#define STF_PACK_MINB 5
#endif
constexpr int PACK_UNROLL = STF_PACK_UNROLL;
__global__ void __launch_bounds__(256, STF_PACK_MINB) k_pack_build_small(const uint8_t* __restrict__ payload, const uint64_t* __restrict__ poff, const int32_t* __restrict__ list, int nlist,
                                                          uint32_t* __restrict__ words, LevelMeta* __restrict__ lm, int L, int* __restrict__ err, int grp) {
    __shared__ uint32_t sbuf[8][2 * RB_HALF];
    const unsigned FULLM = 0xffffffffu;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    uint32_t hi_acc = 0, zero_acc = 0;
    for (int g = warp; g * grp < nlist; g += nwarps) { // grp (4..32) objects per warp: enough groups to fill the GPU
        int my_i = g * grp + lane;
        int my_o = (lane < grp && my_i < nlist) ? list[my_i] : -1;
        // the level-1 meta and the payload offset of EVERY object of the group in one round trip (lane k: object k): the payload
        // loads of an object then depend on nothing that is still in flight, and the load of its other metas overlaps them
        LevelMeta my_m1 = {}; uint64_t my_poff = 0;
        if (my_o >= 0) { my_m1 = lm[(size_t)my_o * L]; my_poff = poff[my_o]; }
        // ---- phase A
        for (int k = 0; k < grp; k++) {
            int o = __shfl_sync(FULLM, my_o, k);
            if (o < 0) break;
            LevelMeta mine = {};
            if (lane >= 1 && lane < L) mine = lm[(size_t)o * L + lane];
            const LevelMeta m1 = shfl_meta(my_m1, k);
            if (lane == 0) mine = m1;
            int kh = lm_kh(m1), wpc = wpc_of(kh), nw = wpc * lm_kw(m1);
            const uint8_t* base = payload + __shfl_sync(FULLM, my_poff, k);
            uint32_t d = lm_dflt(m1);
            const uint32_t inv = 0xFFFFFFFFu / (uint32_t)wpc + 1u; // j / wpc == umulhi(j, inv) while j * wpc < 2^32
            const uint32_t tm = tail_mask(kh), dwd = dflt_word(d);
#pragma unroll PACK_UNROLL
            for (int j = lane; j < nw; j += 32) { // independent iterations: unrolled so that the loads of several words are in flight together
                uint32_t c = wpc > 1 ? __umulhi((uint32_t)j, inv) : (uint32_t)j, w = (uint32_t)j - c * (uint32_t)wpc; // (inv overflows for wpc == 1)
                uint32_t out = pack16(base + (c * (uint32_t)kh + 16u * w), (int)w == wpc - 1 ? tm : 0xFFFFFFFFu, dwd, hi_acc, zero_acc);
                words[m1.off + j] = out;
                sbuf[wib][j] = out;
            }
            __syncwarp();
            mine = warp_rebuild<false, false, true>(words, mine, L, 1, lane, sbuf[wib], tail_level(kh, lm_kw(m1), L));
            if (lane >= 1 && lane < L) lm[(size_t)o * L + lane] = mine;
            __syncwarp();
        }
        // ---- phase B: lane k finishes object k
        if (my_o >= 0) thread_tail_build<false>(words, lm + (size_t)my_o * L, L, 1);
        __syncwarp();
    }
    if ((hi_acc & 0xFCFCFCFCu) | zero_acc) atomicOr(err, 1);
}

// levels from+1..L of one object by one CTA: level `from` is read from global memory (L2-hot), every later level from
// the shared-memory copy of the level just produced (A holds levels from+1, from+3, .., B the others)
template <int THREADS> __device__ __forceinline__ void cta_build_levels(uint32_t* __restrict__ words, const LevelMeta* sm, int L, int from, uint32_t* A, uint32_t* B) {
    const int tid = threadIdx.x;
    const uint32_t* src = words + sm[from - 1].off;
    uint32_t* dst = A;
    for (int l = from + 1; l <= L; l++) {
        LevelMeta ch = sm[l - 2], p = sm[l - 1];
        int er = ch.rs - 2 * p.rs, ec = ch.cs - 2 * p.cs;
        int pw_ = wpc_of(lm_kh(p)), pn = pw_ * lm_kw(p);
        for (int j = tid; j < pn; j += THREADS) {
            int pc = j / pw_, pw = j - pc * pw_;
            uint32_t v = build_word<false>(src, ch, er, ec, pc, pw);
            words[p.off + j] = v;
            dst[j] = v;
        }
        __syncthreads();
        src = dst; dst = (dst == A) ? B : A;
    }
}
// ---- fused pack + build, mid-size objects: one CTA per object (THREADS = 128 for a few KB, 512 for mask-sized ones).
// Level 1 is packed straight from global memory into global memory (4 words per thread in flight; it is re-read once,
// L2-hot, to form level 2); level 2 and everything above are staged in shared memory, sized by the host to the largest
// listed object (a_words for levels 2,4,.., b_words for levels 3,5,..) so that several CTAs share an SM.
// (A persistent variant that streamed the payload through shared memory with TMA bulk copies — cp.async.bulk + mbarrier,
// 3 x 32 KB in flight per SM — measured 1.5x SLOWER on B200: one 512-thread CTA per SM cannot issue the pack arithmetic
// fast enough; the kernel is bound by instruction issue, not by the load path.)
template <int THREADS> __global__ void __launch_bounds__(THREADS) k_pack_build_mid(const uint8_t* __restrict__ payload, const uint64_t* __restrict__ poff, const int32_t* __restrict__ list,
                                                                                   uint32_t* __restrict__ words, LevelMeta* __restrict__ lm, int L, int a_words, int* __restrict__ err) {
    extern __shared__ uint32_t pbm[]; // A: a_words, B: the rest
    __shared__ LevelMeta sm[STF_MAX_LEVELS];
    uint32_t* A = pbm; uint32_t* B = pbm + a_words;
    const int o = list[blockIdx.x], tid = threadIdx.x;
    if (tid < L) sm[tid] = lm[(size_t)o * L + tid];
    __syncthreads();
    if (tid == 0) for (int l = 1; l < L; l++) { sm[l].rs = sm[l - 1].rs / 2; sm[l].cs = sm[l - 1].cs / 2; } // qtrees.jl:225-226
    LevelMeta m1 = sm[0];
    const int kh = lm_kh(m1), wpc = wpc_of(kh), nw = wpc * lm_kw(m1);
    const uint8_t* base = payload + poff[o];
    const uint32_t d = lm_dflt(m1);
    uint32_t hi_acc = 0, zero_acc = 0;
    uint32_t* out = words + m1.off;
    { // words j = tid, tid + THREADS, ...: (column, word) and the byte offset advance incrementally, all in 32 bits
        const uint32_t tm = tail_mask(kh), dwd = dflt_word(d);
        int c0 = tid / wpc, w = tid - c0 * wpc;
        uint32_t off = (uint32_t)c0 * (uint32_t)kh + 16u * (uint32_t)w;
        const int dc = THREADS / wpc, dwi = THREADS - dc * wpc;
        const uint32_t doff = (uint32_t)dc * (uint32_t)kh + 16u * (uint32_t)dwi, carry = (uint32_t)kh - 16u * (uint32_t)wpc;
#pragma unroll 4
        for (int j = tid; j < nw; j += THREADS) {
            out[j] = pack16(base + off, w == wpc - 1 ? tm : 0xFFFFFFFFu, dwd, hi_acc, zero_acc);
            w += dwi; off += doff;
            if (w >= wpc) { w -= wpc; off += carry; }
        }
    }
    if ((hi_acc & 0xFCFCFCFCu) | zero_acc) atomicOr(err, 1);
    __syncthreads();
    if (tid >= 1 && tid < L) lm[(size_t)o * L + tid] = sm[tid];
    cta_build_levels<THREADS>(words, sm, L, 1, A, B);
}
#define PBM_A 16384   // largest level 2 (words) a CTA stages; beyond that the object is built by wide launches
#define PBM_B 4608

// ---- warp-per-object build: worklist entry = (object, from_level): levels from+1..L.  32 entries per warp: the levels
// with real width by the whole warp (warp_rebuild), the tail levels by one thread per entry (thread_tail_build) ----------
// Fused shift!/setshift! (qtrees.jl:267-287): with `srs` != NULL entry i is (labels[i] or i, level): levels <= level move by
// s*2^(level-l) (add) or are set to it, then levels > level are rebuilt — one launch instead of shift + rebuild.
__global__ void k_build_objects(uint32_t* __restrict__ words, LevelMeta* __restrict__ lm, int L,
                                const int2* __restrict__ work, int nwork, const ObjMeta* __restrict__ om, int skip_big, int grp,
                                const int32_t* __restrict__ labels = nullptr, const int32_t* __restrict__ srs = nullptr, const int32_t* __restrict__ scs = nullptr,
                                int level = 0, int add = 0, int* __restrict__ built_from = nullptr, int skip_same = 0,
                                int check_nobj = 0, int* __restrict__ err = nullptr) {
    pdl_sync(); // programmatic dependent launch: scheduled while the previous kernel drains; nothing global before this
    __shared__ uint32_t sbuf[8][2 * RB_HALF]; // blockDim.x == 256
    const unsigned FULLM = 0xffffffffu;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int g = warp; g * grp < nwork; g += nwarps) { // grp (1..32) entries per warp: few entries -> more warps, shorter chains
        int my_i = g * grp + lane;
        int2 mine_w = make_int2(-1, 0); int my_s1 = 0, my_s2 = 0;
        if (lane < grp && my_i < nwork) {
            if (srs) { mine_w = make_int2(labels ? labels[my_i] - 1 : my_i, level); my_s1 = srs[my_i]; my_s2 = scs[my_i]; }
            else mine_w = work[my_i];
            // device-resident label lists were not seen by the host: out-of-range labels and mask-sized objects (rebuilt by
            // host-driven wide launches) are rejected here, the entry is dropped and the error surfaces at the next sync
            if (check_nobj && ((unsigned)mine_w.x >= (unsigned)check_nobj || om[mine_w.x].big)) { atomicOr(err, 8); mine_w.x = -1; }
        }
        bool big_mine = mine_w.x >= 0 && skip_big && om[mine_w.x].big;
        if (mine_w.x >= 0 && !srs && (big_mine || mine_w.y >= L)) mine_w.x = -1;
        for (int k = 0; k < grp; k++) {
            int o = __shfl_sync(FULLM, mine_w.x, k), from = __shfl_sync(FULLM, mine_w.y, k);
            if (o < 0) continue;
            LevelMeta mine = {};
            if (lane < L) mine = lm[(size_t)o * L + lane];
            if (srs) { // the shift itself: level l <= `level` moves by s * 2^(level - l)
                int s1 = __shfl_sync(FULLM, my_s1, k), s2 = __shfl_sync(FULLM, my_s2, k);
                bool bigk = __shfl_sync(FULLM, (int)big_mine, k) != 0;
                if (skip_same && !add && !bigk && built_from[o] < from) {
                    // the tree already sits there: its last rebuild started at or below `level` (so the rasters above are
                    // the chain of that level's raster), every level at or below `level` has the requested shift and the
                    // levels above carry its ÷2 chain: the rebuild would write what is there
                    int prs = __shfl_up_sync(FULLM, mine.rs, 1), pcs = __shfl_up_sync(FULLM, mine.cs, 1);
                    int f = 1 << max(from - 1 - lane, 0);
                    bool same = lane >= L || (lane < from ? (mine.rs == s1 * f && mine.cs == s2 * f) : (mine.rs == prs / 2 && mine.cs == pcs / 2));
                    if (__all_sync(FULLM, same)) { if (lane == k) mine_w.x = -1; continue; }
                }
                if (lane < from) {
                    int f = 1 << (from - 1 - lane);
                    if (add) { mine.rs += s1 * f; mine.cs += s2 * f; } else { mine.rs = s1 * f; mine.cs = s2 * f; }
                    lm[(size_t)o * L + lane] = mine;
                }
                if (built_from && lane == 0) built_from[o] = from - 1;
                if (bigk || from >= L) { __syncwarp(); continue; } // mask-sized objects are rebuilt by wide launches
            } else if (built_from && lane == 0) built_from[o] = from - 1;
            LevelMeta mf = shfl_meta(mine, from - 1);
            int T = grp >= 4 ? tail_level_from(lm_kh(mf), lm_kw(mf), L, from) : L; // a thread tail pays off from ~4 objects per warp
            mine = warp_rebuild<false>(words, mine, L, from, lane, sbuf[threadIdx.x >> 5], T);
            if (lane >= from && lane < L) lm[(size_t)o * L + lane] = mine;
            __syncwarp();
        }
        if (grp >= 4 && mine_w.x >= 0 && !(srs && (big_mine || mine_w.y >= L))) thread_tail_build<false>(words, lm + (size_t)mine_w.x * L, L, mine_w.y);
        __syncwarp();
    }
}

// ---- wide build of ONE level of ONE (big) object: grid-stride over parent words.  The level's shift (child ÷ 2,
// qtrees.jl:225-226) is derived by every thread from the child meta and stored by one thread: nobody in this launch
// reads it, the next level's launch does.
__global__ void k_build_level_wide(uint32_t* __restrict__ words, LevelMeta* __restrict__ lm, int L, int o, int l) {
    LevelMeta ch = lm[(size_t)o * L + (l - 2)];
    LevelMeta p = lm[(size_t)o * L + (l - 1)];
    int rs = ch.rs / 2, cs = ch.cs / 2;
    int er = ch.rs - 2 * rs, ec = ch.cs - 2 * cs;
    int wpc = wpc_of(lm_kh(p));
    int64_t nw = (int64_t)wpc * lm_kw(p);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (int64_t)gridDim.x * blockDim.x) {
        int pc = (int)(i / wpc), pw = (int)(i - (int64_t)pc * wpc);
        words[p.off + i] = build_word<false>(words + ch.off, ch, er, ec, pc, pw);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { lm[(size_t)o * L + (l - 1)].rs = rs; lm[(size_t)o * L + (l - 1)].cs = cs; }
}
// once a level of a big object fits a CTA's shared memory, ONE launch finishes the pyramid: levels from+1..L
__global__ void __launch_bounds__(512) k_build_tail_cta(uint32_t* __restrict__ words, LevelMeta* __restrict__ lm, int L, int o, int from, int a_words) {
    extern __shared__ uint32_t pbm[];
    __shared__ LevelMeta sm[STF_MAX_LEVELS];
    const int tid = threadIdx.x;
    if (tid < L) sm[tid] = lm[(size_t)o * L + tid];
    __syncthreads();
    if (tid == 0) for (int l = from; l < L; l++) { sm[l].rs = sm[l - 1].rs / 2; sm[l].cs = sm[l - 1].cs / 2; }
    __syncthreads();
    if (tid >= from && tid < L) lm[(size_t)o * L + tid] = sm[tid];
    cta_build_levels<512>(words, sm, L, from, pbm, pbm + a_words);
}
// ---- overlap!(ground, trees) at level 1 (qtree_functions.jl:452-497), all trees at once.  Elementwise rule
// overlap(p1, p2) = ((p1 ⊻ EMPTY) | (p2 ⊻ EMPTY)) ⊻ EMPTY on the 2-bit codes = OR of the "contains full" bits, AND of
// the "contains empty" bits: two commutative atomics per touched ground word, so the result does not depend on the
// order of the trees.  One warp per listed tree; writes are clipped to the ground kernel (outside it the ground reads
// `default` = FULL for a mask and _overlap! never descends).  The upper ground levels are rebuilt afterwards.
__global__ void __launch_bounds__(256) k_overlap_level1(const uint32_t* __restrict__ words, const LevelMeta* __restrict__ lm, int L, const int32_t* __restrict__ list, int nlist,
                                                        uint32_t* __restrict__ gwords, LevelMeta g) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int gkh = lm_kh(g), gkw = lm_kw(g), gwpc = wpc_of(gkh);
    for (int i = warp; i < nlist; i += nwarps) {
        int o = list[i];
        LevelMeta m = lm[(size_t)o * L];
        int kh = lm_kh(m), kw = lm_kw(m), wpc = wpc_of(kh), nw = wpc * kw;
        int dr = m.rs - g.rs, dc = m.cs - g.cs;
        for (int j = lane; j < nw; j += 32) {
            int c = j / wpc, w = j - c * wpc;
            int gc = dc + c;
            if ((unsigned)gc >= (unsigned)gkw) continue;
            int gr0 = dr + 16 * w;
            int lo = max(0, -gr0), hi = min(min(16, kh - 16 * w), gkh - gr0);
            if (hi <= lo) continue;
            uint32_t vm = (hi >= 16 ? 0xFFFFFFFFu : ((1u << (2 * hi)) - 1u)) & ~((1u << (2 * lo)) - 1u);
            uint32_t x = words[m.off + j];
            uint64_t H = (uint64_t)(x & 0xAAAAAAAAu & vm), Z = (uint64_t)(~x & 0x55555555u & vm);
            int gw0 = gr0 >> 4, sh = (gr0 & 15) * 2;
            H <<= sh; Z <<= sh;
            uint32_t* gp = gwords + g.off + (size_t)gc * gwpc;
            uint32_t h0 = (uint32_t)H, h1 = (uint32_t)(H >> 32), z0 = (uint32_t)Z, z1 = (uint32_t)(Z >> 32);
            if (gw0 >= 0 && gw0 < gwpc) { if (h0) atomicOr(gp + gw0, h0); if (z0) atomicAnd(gp + gw0, ~z0); }
            if (gw0 + 1 >= 0 && gw0 + 1 < gwpc) { if (h1) atomicOr(gp + gw0 + 1, h1); if (z1) atomicAnd(gp + gw0 + 1, ~z1); }
        }
    }
}

// ---- work list of stf_build: (object, from-level) entries; shift!/setshift! (qtrees.jl:267-287) are fused into k_build_objects
__global__ void k_fill_worklist(int n, const int32_t* __restrict__ labels, int from, int2* __restrict__ work) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) work[i] = make_int2(labels ? labels[i] - 1 : i, from);
}

// ---- unpack one level to u8 (download / tests) ---------------------------------------------------
__global__ void k_unpack_level(const uint32_t* __restrict__ words, const LevelMeta* __restrict__ lm, int L, int o, int l, uint8_t* __restrict__ out) {
    LevelMeta m = lm[(size_t)o * L + (l - 1)];
    int kh = lm_kh(m), kw = lm_kw(m), wpc = wpc_of(kh);
    int64_t n = (int64_t)kh * kw;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i / kh), r = (int)(i - (int64_t)c * kh);
        out[i] = (uint8_t)((words[m.off + (size_t)c * wpc + (r >> 4)] >> ((r & 15) << 1)) & 3u);
    }
}

// every level of the listed trees in ONE launch (stf_download_trees): block x handles (tree x / L, level x % L + 1);
// boff[x] = byte offset of that level in `out` (host-computed prefix sums of kh*kw)
__global__ void k_unpack_trees(const uint32_t* __restrict__ words, const LevelMeta* __restrict__ lm, int L, const int32_t* __restrict__ labels,
                               const uint64_t* __restrict__ boff, uint8_t* __restrict__ out) {
    int i = blockIdx.x / L, l = blockIdx.x % L;
    int o = labels ? labels[i] - 1 : i;
    LevelMeta m = lm[(size_t)o * L + l];
    int kh = lm_kh(m), kw = lm_kw(m), wpc = wpc_of(kh);
    uint8_t* dst = out + boff[blockIdx.x];
    for (int j = threadIdx.x; j < kh * kw; j += blockDim.x) {
        int c = j / kh, r = j - c * kh;
        dst[j] = (uint8_t)((words[m.off + (size_t)c * wpc + (r >> 4)] >> ((r & 15) << 1)) & 3u);
    }
}

// t[l,a,b] for n queries (stf_read_nodes) and grad2d (fit_helper.jl:51-66)
__global__ void k_read_nodes(const uint32_t* __restrict__ words, const LevelMeta* __restrict__ lm, int L, int n,
                             const int32_t* __restrict__ labels, const int32_t* __restrict__ l, const int32_t* __restrict__ a,
                             const int32_t* __restrict__ b, uint8_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    LevelMeta m = lm[(size_t)(labels[i] - 1) * L + (l[i] - 1)];
    out[i] = (uint8_t)node_code<false>(words, m, a[i], b[i]);
}
template <bool CG> __device__ __forceinline__ int2 grad2d_dev(const uint32_t* __restrict__ words, const LevelMeta& m, int a, int b) {
#define STF_D(r, c) decode2(node_code<CG>(words, m, (r), (c)))
    int diag = -STF_D(a - 1, b - 1) + STF_D(a + 1, b + 1);
    int cdiag = -STF_D(a - 1, b + 1) + STF_D(a + 1, b - 1);
    int gh = diag + cdiag - STF_D(a - 1, b) + STF_D(a + 1, b);
    int gw = diag - cdiag - STF_D(a, b - 1) + STF_D(a, b + 1);
#undef STF_D
    return make_int2(gh, gw);
}
__global__ void k_grad2d(const uint32_t* __restrict__ words, const LevelMeta* __restrict__ lm, int L, int n,
                         const int32_t* __restrict__ labels, const int32_t* __restrict__ l, const int32_t* __restrict__ a,
                         const int32_t* __restrict__ b, int32_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    LevelMeta m = lm[(size_t)(labels[i] - 1) * L + (l[i] - 1)];
    int2 g = grad2d_dev<false>(words, m, a[i], b[i]);
    out[2 * i] = g.x; out[2 * i + 1] = g.y;
}
