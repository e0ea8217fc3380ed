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

#define STF_MAX_LEVELS 17          // canvas <= 65536; kh,kw <= 32767 (15-bit fields)
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

// 32 consecutive child fields of one child column, starting at 0-based row 32*pw - er
template <bool CG> __device__ __forceinline__ uint64_t child_fields(const uint32_t* __restrict__ words, const LevelMeta& ch, int col, int pw, int er) {
    uint32_t dw = dflt_word(lm_dflt(ch));
    int wpc = wpc_of(lm_kh(ch));
    if ((unsigned)col >= (unsigned)lm_kw(ch)) return ((uint64_t)dw << 32) | dw;
    const uint32_t* base = words + ch.off + (size_t)col * wpc;
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
template <bool CG> __device__ __forceinline__ uint32_t build_word(const uint32_t* __restrict__ words, const LevelMeta& ch, int er, int ec, int pc, int pw) {
    int c0 = 2 * pc - ec;
    uint64_t y = child_fields<CG>(words, ch, c0, pw, er) | child_fields<CG>(words, ch, c0 + 1, pw, er);
    uint64_t z = y | (y >> 2);          // even fields hold the OR of each row pair
    z &= 0x3333333333333333ull;         // compress the even 2-bit fields
    z = (z | (z >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    z = (z | (z >> 4)) & 0x00FF00FF00FF00FFull;
    z = (z | (z >> 8)) & 0x0000FFFF0000FFFFull;
    z = (z | (z >> 16)) & 0x00000000FFFFFFFFull;
    return (uint32_t)z;
}

// buildqtree!(t, from+1) for one object by one warp (qtrees.jl:221-237): levels from+1..L are recomputed
// from level `from`, each level's shift = child shift ÷ 2.
template <bool CG> __device__ __forceinline__ void rebuild_levels_warp(uint32_t* __restrict__ words, LevelMeta* __restrict__ lm, int o, int L, int from, int lane) {
    LevelMeta ch = ldmeta<CG>(lm + (size_t)o * L + (from - 1));
    for (int l = from + 1; l <= L; l++) {
        LevelMeta* pp = lm + (size_t)o * L + (l - 1);
        LevelMeta p = ldmeta<CG>(pp);
        int rs = ch.rs / 2, cs = ch.cs / 2;
        int er = ch.rs - 2 * rs, ec = ch.cs - 2 * cs;
        p.rs = rs; p.cs = cs;
        if (lane == 0) {
            uint4 v = make_uint4(p.off, (uint32_t)rs, (uint32_t)cs, p.dims);
            if (CG) __stcg(reinterpret_cast<uint4*>(pp), v); else *reinterpret_cast<uint4*>(pp) = v;
        }
        int wpc = wpc_of(lm_kh(p));
        int nw = wpc * lm_kw(p);
        for (int i = lane; i < nw; i += 32) {
            int pc = i / wpc, pw = i - pc * wpc;
            stw<CG>(words + p.off + i, build_word<CG>(words, ch, er, ec, pc, pw));
        }
        __syncwarp();
        ch = p;
    }
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
