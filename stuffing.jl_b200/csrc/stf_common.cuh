// stf_common.cuh — data layout in HBM and the device helpers every kernel shares.
//
// Layout:
//  * node codes EMPTY=1/FULL=2/MIX=3 (qtrees.jl:56) are stored 2 bits per node, 16 nodes per
//    32-bit word, packed along the ROW index `a` (Julia's fast, column-major index); one level of
//    one object is kw columns of wpc = ceil(kh/16) words.  Fields past kh in the last word of a
//    column hold the object's `default`, so whole-word operations never see garbage.
//  * no physical padding: PaddedMat's 1+2 pad rows/cols (qtrees.jl:117-128) only ever hold
//    `default`, so a read outside [1..kh]x[1..kw] returns `default` (safe-mode getindex, :156-162).
//  * one 16-byte LevelMeta per (object, level) carries everything a node read needs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define STF_MAX_LEVELS 16          // canvas <= 32768; kh,kw <= 32767 (15-bit fields); one half-warp of LevelMetas
#define STF_CELL_W 22              // bits per cell coordinate in a sort key (biased by 2^21)
#define STF_CELL_BIAS (1 << (STF_CELL_W - 1))
#define STF_KEY_INVALID 0xFFFFFFFFFFFFFFFFull

struct __align__(16) LevelMeta {
    uint32_t off;   // word offset of this level in the node store
    int32_t rs;     // rshift   (PaddedMat.rshift, qtrees.jl:112)
    int32_t cs;     // cshift
    uint32_t dims;  // kh | kw<<15 | default<<30
};
__host__ __device__ __forceinline__ int lm_kh(const LevelMeta& m) { return (int)(m.dims & 0x7fffu); }
__host__ __device__ __forceinline__ int lm_kw(const LevelMeta& m) { return (int)((m.dims >> 15) & 0x7fffu); }
__host__ __device__ __forceinline__ uint32_t lm_dflt(const LevelMeta& m) { return m.dims >> 30; }
__host__ __device__ __forceinline__ int wpc_of(int kh) { return (kh + 15) >> 4; }
__host__ __device__ __forceinline__ uint32_t dflt_word(uint32_t d) { return d * 0x55555555u; }

struct ObjMeta {
    int32_t kh1, kw1;   // level-1 kernel size (kernelsize(t[1]))
    int32_t scene;
    uint8_t dflt, is_mask, big, pad;
};

// internal 16-byte item: (i1, i2 | l<<26, a, b)
struct __align__(16) Item16 { int32_t i1, i2l, a, b; };
__host__ __device__ __forceinline__ Item16 make_item(int i1, int i2, int l, int a, int b) { Item16 x = { i1, i2 | (l << 26), a, b }; return x; }
__host__ __device__ __forceinline__ int item_i2(const Item16& x) { return x.i2l & 0x3ffffff; }
__host__ __device__ __forceinline__ int item_l(const Item16& x) { return (int)((uint32_t)x.i2l >> 26); }

#ifdef __CUDACC__
template <bool CG> __device__ __forceinline__ uint32_t ldw(const uint32_t* p) { return CG ? __ldcg(p) : *p; }
template <bool CG> __device__ __forceinline__ void stw(uint32_t* p, uint32_t v) { if (CG) __stcg(p, v); else *p = v; }
template <bool CG> __device__ __forceinline__ LevelMeta ldmeta(const LevelMeta* p) {
    uint4 v = CG ? __ldcg(reinterpret_cast<const uint4*>(p)) : *reinterpret_cast<const uint4*>(p);
    LevelMeta m; m.off = v.x; m.rs = (int)v.y; m.cs = (int)v.z; m.dims = v.w; return m;
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute may be
// scheduled while its predecessor in the stream is still running.  Every such kernel starts with pdl_sync(): wait until
// the predecessor grid has completed and its writes are visible, THEN let the own successor be scheduled (it parks in
// its wait while this grid runs: its launch latency and prologue are off the critical path).  The order matters:
// triggering before the wait lets a chain A -> B -> C schedule C while A still runs, and on B200 C's wait was then
// observed to return before A's writes were complete (a second call on one context lost its candidate list).
// Launched without the attribute, both instructions are no-ops.
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// t[l][a,b] — PaddedMat getindex (qtrees.jl:156-162) on the packed store
template <bool CG> __device__ __forceinline__ uint32_t node_code(const uint32_t* __restrict__ words, const LevelMeta& m, int a, int b) {
    int r = a - m.rs - 1, c = b - m.cs - 1;
    int kh = lm_kh(m);
    if ((unsigned)r < (unsigned)kh && (unsigned)c < (unsigned)lm_kw(m)) {
        uint32_t w = ldw<CG>(words + m.off + (size_t)c * wpc_of(kh) + (r >> 4));
        return (w >> ((r & 15) << 1)) & 3u;
    }
    return lm_dflt(m);
}

// OR of adjacent 2-bit field pairs, then keep the even fields: 16 fields (32 bits) -> 8 fields (16 bits).
// Done on 32-bit halves so that every step is one shift + one LOP3 and the two halves overlap (ILP).
__device__ __forceinline__ uint32_t compress_half(uint32_t h) {
    uint32_t x = (h | (h >> 2)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}
__device__ __forceinline__ uint32_t compress_pairs(uint64_t y) {
    return compress_half((uint32_t)y) | (compress_half((uint32_t)(y >> 32)) << 16);
}

// 32 consecutive child fields of one child column, starting at 0-based row 32*pw - er.
// `lvl` points at the first word of the child level (global memory, or a shared-memory copy of it)
template <bool CG> __device__ __forceinline__ uint64_t child_fields(const uint32_t* __restrict__ lvl, const LevelMeta& ch, int col, int pw, int er) {
    uint32_t dw = dflt_word(lm_dflt(ch));
    int wpc = wpc_of(lm_kh(ch));
    if ((unsigned)col >= (unsigned)lm_kw(ch)) return ((uint64_t)dw << 32) | dw;
    const uint32_t* base = lvl + (size_t)col * wpc;
    int k0 = 2 * pw;
    uint32_t w0 = k0 < wpc ? ldw<CG>(base + k0) : dw;
    uint32_t w1 = k0 + 1 < wpc ? ldw<CG>(base + k0 + 1) : dw;
    uint64_t x = ((uint64_t)w1 << 32) | w0;
    if (er > 0) {
        uint32_t wm = (k0 - 1 >= 0 && k0 - 1 < wpc) ? ldw<CG>(base + k0 - 1) : dw;
        x = (x << 2) | (wm >> 30);
    } else if (er < 0) {
        uint32_t wp = k0 + 2 < wpc ? ldw<CG>(base + k0 + 2) : dw;
        x = (x >> 2) | ((uint64_t)(wp & 3u) << 62);
    }
    return x;
}

// one word (16 nodes) of a parent level: qcode (qtrees.jl:50-53) on 2-bit fields.
// er/ec = child shift - 2*parent shift in {-1,0,1} (÷2 truncates toward zero, qtrees.jl:225-226)
template <bool CG> __device__ __forceinline__ uint32_t build_word(const uint32_t* __restrict__ lvl, const LevelMeta& ch, int er, int ec, int pc, int pw) {
    int c0 = 2 * pc - ec;
    uint64_t y = child_fields<CG>(lvl, ch, c0, pw, er) | child_fields<CG>(lvl, ch, c0 + 1, pw, er);
    return compress_pairs(y);
}

__device__ __forceinline__ LevelMeta shfl_meta(const LevelMeta& m, int src) {
    LevelMeta r;
    r.off = __shfl_sync(0xffffffffu, m.off, src); r.rs = __shfl_sync(0xffffffffu, m.rs, src);
    r.cs = __shfl_sync(0xffffffffu, m.cs, src); r.dims = __shfl_sync(0xffffffffu, m.dims, src);
    return r;
}
template <bool CG> __device__ __forceinline__ void stmeta(LevelMeta* p, const LevelMeta& m) {
    uint4 v = make_uint4(m.off, (uint32_t)m.rs, (uint32_t)m.cs, m.dims);
    if (CG) __stcg(reinterpret_cast<uint4*>(p), v); else *reinterpret_cast<uint4*>(p) = v;
}

// buildqtree!(t, from+1) for one object by one warp (qtrees.jl:221-237): levels from+1..L are recomputed
// from level `from`, each level's shift = child shift ÷ 2.
// Latency design: lane i holds the LevelMeta of level i+1 (`mine`, loaded by the caller in ONE round trip);
// the first rebuilt level reads its children from global memory, every later level reads them from a
// per-warp shared-memory copy of the level just produced (levels shrink 4x, so they fit after one step);
// the global stores are fire-and-forget.  Returns the lane's meta with the new shifts (lanes >= from).
#define RB_HALF 384   // words per shared-memory level buffer (two per warp)
// KEEP: the per-level shifts in `mine` are already final (lazy batchsteps tracks them); only the rasters are rebuilt.
// FIRST_SMEM: level `from` has already been staged in sbuf[0 .. ) by the caller (fused pack + build).
// upto: last level whose raster is built (shifts are still computed for every level).
template <bool CG, bool KEEP = false, bool FIRST_SMEM = false> __device__ __forceinline__ LevelMeta warp_rebuild(uint32_t* __restrict__ words, LevelMeta mine, int L, int from, int lane, uint32_t* __restrict__ sbuf, int upto = STF_MAX_LEVELS) {
    const unsigned FULLM = 0xffffffffu;
    // shifts of the rebuilt levels in closed form: repeated `÷ 2` (toward zero) composes to `÷ 2^k`
    int brs = __shfl_sync(FULLM, mine.rs, from - 1), bcs = __shfl_sync(FULLM, mine.cs, from - 1);
    if (!KEEP && lane >= from && lane < L) { // x / 2^sh toward zero without an integer division
        int sh = lane + 1 - from;
        mine.rs = brs >= 0 ? (brs >> sh) : -((-brs) >> sh); mine.cs = bcs >= 0 ? (bcs >> sh) : -((-bcs) >> sh);
    }
    // constants of "my" level seen as a parent; its child level is held by lane-1
    int crs = __shfl_up_sync(FULLM, mine.rs, 1), ccs = __shfl_up_sync(FULLM, mine.cs, 1);
    uint32_t cdims = __shfl_up_sync(FULLM, mine.dims, 1), coff = __shfl_up_sync(FULLM, mine.off, 1);
    int my_er = crs - 2 * mine.rs, my_ec = ccs - 2 * mine.cs;
    bool in_smem = FIRST_SMEM; int cur = 0;
    bool in_reg = false; uint32_t regw = 0; // once a level is one word per column and <= 32 columns, lane j keeps column j
    const int last = min(L, upto);
    for (int l = from + 1; l <= last; l++) {
        int s = l - 1;
        uint32_t poff = __shfl_sync(FULLM, mine.off, s), pdims = __shfl_sync(FULLM, mine.dims, s);
        LevelMeta ch; ch.off = __shfl_sync(FULLM, coff, s); ch.dims = __shfl_sync(FULLM, cdims, s); ch.rs = 0; ch.cs = 0;
        int er = __shfl_sync(FULLM, my_er, s), ec = __shfl_sync(FULLM, my_ec, s);
        int kwp = (int)((pdims >> 15) & 0x7fffu), wpc = wpc_of((int)(pdims & 0x7fffu));
        int wpcc = wpc_of(lm_kh(ch)), kwc = lm_kw(ch);
        if (in_reg) {
            // register path: the child level lives in `regw` (wpcc == 1, kwc <= 32), so kh_p <= 9 and kw_p <= 17;
            // two shuffles fetch the child columns, no memory on the dependent chain, the store is fire-and-forget
            uint32_t d = lm_dflt(ch), dw = dflt_word(d);
            int c0 = 2 * lane - ec;
            uint32_t w0 = __shfl_sync(FULLM, regw, c0 & 31), w1 = __shfl_sync(FULLM, regw, (c0 + 1) & 31);
            if ((unsigned)c0 >= (unsigned)kwc) w0 = dw;
            if ((unsigned)(c0 + 1) >= (unsigned)kwc) w1 = dw;
            uint64_t y = ((uint64_t)dw << 32) | (w0 | w1);
            if (er > 0) y = (y << 2) | d; else if (er < 0) y = (y >> 2) | ((uint64_t)d << 62);
            regw = compress_pairs(y);
            if (lane < kwp) stw<CG>(words + poff + lane, regw);
            continue;
        }
        int nw = wpc * kwp;
        bool fits = nw <= RB_HALF;
        uint32_t* dst = sbuf + (cur ^ 1) * RB_HALF;
        const uint32_t* src = in_smem ? sbuf + cur * RB_HALF : words + ch.off;
        if (wpcc <= 2 && wpc == 1 && kwp <= 32) {
            // small level: one lane per parent column, both child columns are at most two words (kh <= 32)
            uint32_t wv = 0;
            if (lane < kwp) {
                uint32_t d = lm_dflt(ch), dw = dflt_word(d);
                uint64_t y = 0;
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    int cc = 2 * lane - ec + t;
                    uint32_t w0 = dw, w1 = dw;
                    if ((unsigned)cc < (unsigned)kwc) {
                        const uint32_t* bp = src + cc * wpcc;
                        w0 = in_smem ? bp[0] : ldw<CG>(bp);
                        if (wpcc == 2) w1 = in_smem ? bp[1] : ldw<CG>(bp + 1);
                    }
                    y |= ((uint64_t)w1 << 32) | w0;
                }
                if (er > 0) y = (y << 2) | d; else if (er < 0) y = (y >> 2) | ((uint64_t)d << 62);
                wv = compress_pairs(y);
                stw<CG>(words + poff + lane, wv);
            }
            regw = wv; in_reg = true;
            continue;
        }
        for (int i = lane; i < nw; i += 32) {
            int pc = i / wpc, pw = i - pc * wpc;
            uint32_t wv = in_smem ? build_word<false>(src, ch, er, ec, pc, pw) : build_word<CG>(src, ch, er, ec, pc, pw);
            stw<CG>(words + poff + i, wv);
            if (fits) dst[i] = wv;
        }
        __syncwarp();
        in_smem = fits; cur ^= 1;
    }
    return mine;
}
// convenience wrapper: loads and stores the metas itself
template <bool CG> __device__ __forceinline__ void rebuild_levels_warp(uint32_t* __restrict__ words, LevelMeta* __restrict__ lm, int o, int L, int from, int lane, uint32_t* __restrict__ sbuf) {
    LevelMeta mine = {};
    if (lane < L) mine = ldmeta<CG>(lm + (size_t)o * L + lane);
    mine = warp_rebuild<CG>(words, mine, L, from, lane, sbuf);
    if (lane >= from && lane < L) stmeta<CG>(lm + (size_t)o * L + lane, mine);
}


// 32 consecutive 2-bit fields of kernel column c (0-based) from kernel row r0 on (0-based; r0 may be negative or run
// past kh): PaddedMat safe-read semantics, everything outside the kernel is `default`.
template <bool CG> __device__ __forceinline__ uint64_t fields64_at(const uint32_t* __restrict__ words, const LevelMeta& m, int c, int r0) {
    uint32_t dw = dflt_word(lm_dflt(m));
    int wpc = wpc_of(lm_kh(m));
    if ((unsigned)c >= (unsigned)lm_kw(m)) return ((uint64_t)dw << 32) | dw;
    const uint32_t* base = words + m.off + (size_t)c * wpc;
    int w0 = r0 >> 4, off = (r0 & 15) * 2;
    uint32_t x0 = (unsigned)w0 < (unsigned)wpc ? ldw<CG>(base + w0) : dw;
    uint32_t x1 = (unsigned)(w0 + 1) < (unsigned)wpc ? ldw<CG>(base + w0 + 1) : dw;
    uint32_t x2 = (unsigned)(w0 + 2) < (unsigned)wpc ? ldw<CG>(base + w0 + 2) : dw;
    uint64_t lo = ((uint64_t)x1 << 32) | x0;
    if (off) lo = (lo >> off) | ((uint64_t)x2 << (64 - off));
    return lo;
}
__device__ __forceinline__ int decode2(uint32_t c) { return (int)((0x60u >> ((c & 3u) << 1)) & 3u); } // fit_helper.jl:42

// grad2d (fit_helper.jl:51-66) at (l,a,b) of a tree whose levels > m are stale in memory (lazy batchsteps: moves only
// update the per-level shifts, the rasters are materialised once per batch).  The 3x3 window of level l is derived
// from level m through the build rule (qtrees.jl:221-237) with the CURRENT shifts: window origin lo_j = 2*lo_{j+1} - 1,
// lane i = window column i, one 64-bit register = the rows of that column; each level is two shuffles, a pair-OR
// compress and the clip against that level's kernel window (outside -> default).  Needs l - m <= 3 (24 columns).
// `meta_at(j)` returns the LevelMeta of level j (warp-uniform).  Whole warp must call.
// Two halves, so that a caller can issue the loads of several windows before it consumes any of them:
// lazy_window_load = the only memory access (level m's fields, lane = window column), lazy_window_reduce = shuffles only.
template <bool CG> __device__ __forceinline__ uint64_t lazy_window_load(const uint32_t* __restrict__ words, const LevelMeta& mm, int l, int a, int b, int m, int lane) {
    int k = l - m;
    int rlo = a - 1, clo = b - 1;
    for (int i = 0; i < k; i++) { rlo = 2 * rlo - 1; clo = 2 * clo - 1; }
    int n = 3 << k;
    uint64_t x = 0;
    if (lane < n) x = fields64_at<CG>(words, mm, clo + lane - mm.cs - 1, rlo - mm.rs - 1);
    return x;
}
template <class MetaAt> __device__ __forceinline__ int2 lazy_window_reduce(uint64_t x, MetaAt meta_at, int l, int a, int b, int m, int lane) {
    const unsigned FULLM = 0xffffffffu;
    int k = l - m;
    int rlo = a - 1, clo = b - 1;
    for (int i = 0; i < k; i++) { rlo = 2 * rlo - 1; clo = 2 * clo - 1; }
    int n = 3 << k;
    for (int j = m + 1; j <= l; j++) {
        rlo = (rlo + 1) >> 1; clo = (clo + 1) >> 1; n >>= 1; // exact: lo_{j-1} = 2*lo_j - 1
        uint64_t y = __shfl_sync(FULLM, x, (2 * lane) & 31) | __shfl_sync(FULLM, x, (2 * lane + 1) & 31);
        uint32_t p = compress_pairs(y);
        LevelMeta pm = meta_at(j);
        int lo = max(0, pm.rs + 1 - rlo), hi = min(n, pm.rs + lm_kh(pm) + 1 - rlo);
        uint32_t mask = hi > lo ? (((1u << (2 * hi)) - 1u) & ~((1u << (2 * lo)) - 1u)) : 0u;
        int col = clo + lane;
        if (!(col > pm.cs && col <= pm.cs + lm_kw(pm))) mask = 0u;
        x = (uint64_t)((p & mask) | (dflt_word(lm_dflt(pm)) & ~mask));
    }
    int f0 = decode2((uint32_t)x & 3u), f1 = decode2((uint32_t)(x >> 2) & 3u), f2 = decode2((uint32_t)(x >> 4) & 3u);
    int rowdiff = f2 - f0, colsum = f0 + f1 + f2;
    int gh = __shfl_sync(FULLM, rowdiff, 0) + __shfl_sync(FULLM, rowdiff, 1) + __shfl_sync(FULLM, rowdiff, 2);
    int gw = __shfl_sync(FULLM, colsum, 2) - __shfl_sync(FULLM, colsum, 0);
    return make_int2(gh, gw);
}
template <bool CG, class MetaAt> __device__ __forceinline__ int2 grad_lazy(const uint32_t* __restrict__ words, MetaAt meta_at, int l, int a, int b, int m, int lane) {
    LevelMeta mm = meta_at(m);
    uint64_t x = lazy_window_load<CG>(words, mm, l, a, b, m, lane);
    return lazy_window_reduce(x, meta_at, l, a, b, m, lane);
}

// ---- spatial cell keys: scene | level | a+bias | b+bias ----
__device__ __forceinline__ uint64_t cell_key(int scene, int l, int a, int b) {
    return ((uint64_t)(uint32_t)scene << (5 + 2 * STF_CELL_W)) | ((uint64_t)(uint32_t)l << (2 * STF_CELL_W)) |
           ((uint64_t)(uint32_t)(a + STF_CELL_BIAS) << STF_CELL_W) | (uint64_t)(uint32_t)(b + STF_CELL_BIAS);
}
__host__ __device__ __forceinline__ void cell_unkey(uint64_t k, int* l, int* a, int* b) {
    *b = (int)(k & ((1u << STF_CELL_W) - 1)) - STF_CELL_BIAS;
    *a = (int)((k >> STF_CELL_W) & ((1u << STF_CELL_W) - 1)) - STF_CELL_BIAS;
    *l = (int)((k >> (2 * STF_CELL_W)) & 31);
}
#endif

#define CODE_EMPTY 1u
#define CODE_FULL 2u
#define CODE_MIX 3u
