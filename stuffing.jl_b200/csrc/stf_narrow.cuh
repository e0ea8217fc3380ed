// stf_narrow.cuh — subsystem (3), word-parallel part: _collision_randbfs (qtree_functions.jl:21-42, canonical child
// order) for the first THREE levels below the start node as bitwise plane tests on whole blocks of nodes.
//
// The FIFO BFS of the reference visits the tree level by level.  The 2^d x 2^d nodes below the
// start node `at` at depth d (level at.l - d) are fetched as 2-bit fields, 2^d fields per column and tree, AND-ed
// (Q1[ci] & Q2[ci], :30), and split into the two planes of the code: bit 1 "contains full", bit 0 "contains empty":
//     hit  = F & ~E   (code == FULL, :31)        mix = F & E   (code == MIX, :34)
// on every node of the block at once.  Both masks are brought into MORTON order of the child-number paths from `at`
// (digit = child number = (beta << 1) | alpha, alpha = a_hi - a, beta = b_hi - b, most significant digit = first step;
// child 0 is (2a, 2b), qtrees.jl:19-22, so larger coordinates come first): in that order
//   * the BFS order of one level is ascending bit order: the node the sequential BFS returns is the LOWEST set hit bit,
//     the last node it pops the HIGHEST set frontier bit (the (-l, a, b) of a miss, :41);
//   * the children of frontier bit k are bits 4k .. 4k+3 of the next depth: the frontier of depth d, spread by four,
//     masks the block of depth d + 1 — exact for inconsistent pyramids too (negative shifts, raw shift edits), because
//     a node is only ever tested when its parent was MIX in both trees, exactly as the queue does it.
// Depth 1 (4 nodes), 2 (16) and 3 (64 nodes, one 64-bit mask) live in registers of ONE thread: 97 % of the candidate
// pairs of a placed scene end there.  A pair that is still undecided after depth 3 is handed to the frontier kernel
// (k_collide_small) with its depth-3 frontier mask, so nothing is tested twice.
#pragma once
#include "stf_common.cuh"

// 16 consecutive 2-bit fields of kernel column c (0-based) from kernel row r0 on (0-based, may be negative or run past
// kh): PaddedMat safe-read semantics (qtrees.jl:156-162), everything outside the kernel is `default`.  `need` = number
// of fields the caller uses (<= 16): the second word is only loaded when they straddle a word boundary.
template <bool CG> __device__ __forceinline__ uint32_t fields32_at(const uint32_t* __restrict__ words, const LevelMeta& m, int c, int r0, int need) {
    const uint32_t dw = dflt_word(lm_dflt(m));
    const int wpc = wpc_of(lm_kh(m));
    if ((unsigned)c >= (unsigned)lm_kw(m)) return dw;
    const uint32_t* base = words + m.off + (size_t)c * wpc;
    const int w0 = r0 >> 4, off = (r0 & 15) * 2;
    uint32_t x0 = (unsigned)w0 < (unsigned)wpc ? ldw<CG>(base + w0) : dw;
    uint32_t x1 = dw;
    if (off + 2 * need > 32 && (unsigned)(w0 + 1) < (unsigned)wpc) x1 = ldw<CG>(base + w0 + 1);
    return __funnelshift_r(x0, x1, off);
}

// ---- pure bit helpers (host-testable) -------------------------------------------------------------------------------
// odd bits of 16 two-bit fields -> 16 dense bits
__host__ __device__ __forceinline__ uint32_t nw_squeeze_odd(uint32_t x) {
    x = (x >> 1) & 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}
__host__ __device__ __forceinline__ uint32_t nw_brev32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1); x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4); x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
__host__ __device__ __forceinline__ uint64_t nw_brev64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __brevll(x);
#else
    return ((uint64_t)nw_brev32((uint32_t)x) << 32) | nw_brev32((uint32_t)(x >> 32));
#endif
}
// raster masks (bit = column index * 2^d + row index, both ascending in canvas coordinates) -> Morton order of the
// child-number paths.  Reversing the whole mask turns (col, row) into (beta, alpha) = distances from the block's high
// corner; delta swaps then interleave the index bits [beta.. alpha..] -> [beta_hi alpha_hi .. beta_0 alpha_0].
__host__ __device__ __forceinline__ uint32_t nw_morton1(uint32_t r4) { return nw_brev32(r4) >> 28; }
__host__ __device__ __forceinline__ uint32_t nw_morton2(uint32_t r16) {
    uint32_t x = nw_brev32(r16) >> 16;                 // [b1 b0 a1 a0]
    uint32_t t = ((x >> 2) ^ x) & 0x0C0Cu; x ^= t ^ (t << 2); // swap index bits 2 and 1 -> [b1 a1 b0 a0]
    return x;
}
__host__ __device__ __forceinline__ uint64_t nw_morton3(uint64_t r64) {
    uint64_t x = nw_brev64(r64);                       // [b2 b1 b0 a2 a1 a0]
    uint64_t t = ((x >> 12) ^ x) & 0x0000F0F00000F0F0ull; x ^= t ^ (t << 12); // index bits 4 <-> 2: [b2 a2 b0 b1 a1 a0]
    t = ((x >> 4) ^ x) & 0x00F000F000F000F0ull; x ^= t ^ (t << 4);            // index bits 3 <-> 2: [b2 a2 b1 b0 a1 a0]
    t = ((x >> 2) ^ x) & 0x0C0C0C0C0C0C0C0Cull; x ^= t ^ (t << 2);            // index bits 2 <-> 1: [b2 a2 b1 a1 b0 a0]
    return x;
}
// frontier of depth d (Morton) -> mask of its children at depth d + 1: bit k -> bits 4k .. 4k+3
__host__ __device__ __forceinline__ uint32_t nw_children16(uint32_t m4) {
    uint32_t x = m4 & 0xFu;
    x = (x | (x << 6)) & 0x0303u;
    x = (x | (x << 3)) & 0x1111u;
    return x * 0xFu;
}
__host__ __device__ __forceinline__ uint64_t nw_children64(uint32_t m16) {
    uint64_t x = m16 & 0xFFFFu;
    x = (x | (x << 24)) & 0x000000FF000000FFull;
    x = (x | (x << 12)) & 0x000F000F000F000Full;
    x = (x | (x << 6)) & 0x0303030303030303ull;
    x = (x | (x << 3)) & 0x1111111111111111ull;
    return x * 0xFull;
}
// Morton key (2 bits per depth) -> (alpha, beta)
__host__ __device__ __forceinline__ void nw_unkey(uint32_t k, int* alpha, int* beta) {
    uint32_t a = k & 0x55555555u, b = (k >> 1) & 0x55555555u;
    a = (a | (a >> 1)) & 0x33333333u; a = (a | (a >> 2)) & 0x0F0F0F0Fu; a = (a | (a >> 4)) & 0x00FF00FFu; a = (a | (a >> 8)) & 0xFFFFu;
    b = (b | (b >> 1)) & 0x33333333u; b = (b | (b >> 2)) & 0x0F0F0F0Fu; b = (b | (b >> 4)) & 0x00FF00FFu; b = (b | (b >> 8)) & 0xFFFFu;
    *alpha = (int)a; *beta = (int)b;
}

#ifdef __CUDACC__
// the blocks of ONE tree below `at` = (l, a, b): depth 1 (2 columns x 2 fields, 4 bits per column), depth 2 (4 x 4 fields,
// one byte per column), depth 3 (8 x 8 fields, 16 bits per column, two columns per word)
struct NwBlocks { uint32_t p1, p2, p3[4]; };
template <bool CG> __device__ __forceinline__ uint32_t nw_block1(const uint32_t* __restrict__ words, const LevelMeta& m, int a, int b) {
    const int r0 = 2 * a - 1 - m.rs - 1, c0 = 2 * b - 1 - m.cs - 1;
    return (fields32_at<CG>(words, m, c0, r0, 2) & 0xFu) | ((fields32_at<CG>(words, m, c0 + 1, r0, 2) & 0xFu) << 4);
}
template <bool CG> __device__ __forceinline__ uint32_t nw_block2(const uint32_t* __restrict__ words, const LevelMeta& m, int a, int b) {
    const int r0 = 4 * a - 3 - m.rs - 1, c0 = 4 * b - 3 - m.cs - 1;
    uint32_t p = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) p |= (fields32_at<CG>(words, m, c0 + j, r0, 4) & 0xFFu) << (8 * j);
    return p;
}
template <bool CG> __device__ __forceinline__ void nw_block3(const uint32_t* __restrict__ words, const LevelMeta& m, int a, int b, uint32_t p[4]) {
    const int r0 = 8 * a - 7 - m.rs - 1, c0 = 8 * b - 7 - m.cs - 1;
#pragma unroll
    for (int j = 0; j < 4; j++)
        p[j] = (fields32_at<CG>(words, m, c0 + 2 * j, r0, 8) & 0xFFFFu) | (fields32_at<CG>(words, m, c0 + 2 * j + 1, r0, 8) << 16);
}

// result of the word-parallel part for one candidate pair
struct NwResult {
    int l, a, b;        // decided: the BFS result, l >= 0 hit (:32), l < 0 miss (:41)
    int undecided;      // != 0: the frontier of depth 3 (`front`, Morton order, nonzero) goes on in the frontier kernel
    uint64_t front;
    int visited;        // child nodes the sequential BFS tests up to here (the early exit on a hit included)
};
__device__ __forceinline__ int nw_tests_until(uint64_t front, int hitkey) { // BFS tests until it returns child `hitkey`
    int pk = hitkey >> 2;
    return 4 * __popcll(front & ((1ull << pk) - 1ull)) + (hitkey & 3) + 1;
}
// ---- one depth at a time: pa / pb = the blocks of the two trees at that depth, `front` = the frontier of the depth above
// (Morton).  Each returns 0 = decided (R filled), 1 = go one level deeper with the new frontier in *mix.
// depth 1: the four children of `at` = (l, a, b)
__device__ __forceinline__ int nw_depth1(uint32_t pa, uint32_t pb, int l, int a, int b, NwResult& R, uint32_t* mix) {
    uint32_t c = pa & pb;
    uint32_t e = (c << 1) & 0xAAu, f = c & 0xAAu;
    uint32_t hit = nw_morton1(nw_squeeze_odd(f & ~e)); *mix = nw_morton1(nw_squeeze_odd(f & e));
    R.undecided = 0; R.front = 0;
    if (hit) {
        int k = __ffs(hit) - 1;
        R.l = l - 1; R.a = 2 * a - (k & 1); R.b = 2 * b - (k >> 1); R.visited = k + 1;
        return 0;
    }
    R.visited = 4;
    if (*mix == 0 || l - 1 <= 1) { R.l = -l; R.a = a; R.b = b; return 0; } // nothing pushed (:34): the last node popped is `at`
    return 1;
}
__device__ __forceinline__ int nw_depth2(uint32_t pa, uint32_t pb, uint32_t mix1, int l, int a, int b, NwResult& R, uint32_t* mix) {
    uint32_t c = pa & pb;
    uint32_t e = (c << 1) & 0xAAAAAAAAu, f = c & 0xAAAAAAAAu;
    const uint32_t kids = nw_children16(mix1);
    uint32_t hit = nw_morton2(nw_squeeze_odd(f & ~e)) & kids; *mix = nw_morton2(nw_squeeze_odd(f & e)) & kids;
    if (hit) {
        int k = __ffs(hit) - 1, al, be; nw_unkey((uint32_t)k, &al, &be);
        R.l = l - 2; R.a = 4 * a - al; R.b = 4 * b - be; R.visited += nw_tests_until(mix1, k);
        return 0;
    }
    R.visited += 4 * __popc(mix1);
    if (*mix == 0 || l - 2 <= 1) { // miss: the last node popped is the last frontier node of depth 1
        int k = 31 - __clz(mix1);
        R.l = -(l - 1); R.a = 2 * a - (k & 1); R.b = 2 * b - (k >> 1);
        return 0;
    }
    return 1;
}
__device__ __forceinline__ int nw_depth3(const uint32_t pa[4], const uint32_t pb[4], uint32_t mix2, int l, int a, int b, NwResult& R) {
    uint64_t hr = 0, mr = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t cc = pa[j] & pb[j];
        uint32_t ee = (cc << 1) & 0xAAAAAAAAu, ff = cc & 0xAAAAAAAAu;
        hr |= (uint64_t)nw_squeeze_odd(ff & ~ee) << (16 * j);
        mr |= (uint64_t)nw_squeeze_odd(ff & ee) << (16 * j);
    }
    const uint64_t kids = nw_children64(mix2);
    uint64_t hit = hr ? (nw_morton3(hr) & kids) : 0ull, mix3 = mr ? (nw_morton3(mr) & kids) : 0ull;
    if (hit) {
        int k = __ffsll((long long)hit) - 1, al, be; nw_unkey((uint32_t)k, &al, &be);
        R.l = l - 3; R.a = 8 * a - al; R.b = 8 * b - be; R.visited += nw_tests_until(mix2, k);
        return 0;
    }
    R.visited += 4 * __popc(mix2);
    if (mix3 == 0 || l - 3 <= 1) {
        int k = 31 - __clz(mix2), al, be; nw_unkey((uint32_t)k, &al, &be);
        R.l = -(l - 2); R.a = 4 * a - al; R.b = 4 * b - be;
        return 0;
    }
    R.undecided = 1; R.front = mix3; R.l = 0; R.a = 0; R.b = 0;
    return 1;
}
// all three depths in one thread.  LOW: a functor set giving the blocks of tree 1 (precomputed per record in the fused
// candidate kernel); tree 2 is read through (words, lm2) with lm2 = the LevelMeta array of the tree (index level - 1).
template <bool CG, class Low>
__device__ __forceinline__ NwResult nw_collide3(const uint32_t* __restrict__ words, Low low, const LevelMeta* __restrict__ lm2, int l, int a, int b) {
    NwResult R; uint32_t mix1, mix2;
    if (!nw_depth1(low.p1(), nw_block1<CG>(words, ldmeta<CG>(lm2 + (l - 2)), a, b), l, a, b, R, &mix1)) return R;
    if (!nw_depth2(low.p2(), nw_block2<CG>(words, ldmeta<CG>(lm2 + (l - 3)), a, b), mix1, l, a, b, R, &mix2)) return R;
    uint32_t q[4], pl[4]; nw_block3<CG>(words, ldmeta<CG>(lm2 + (l - 4)), a, b, q);
#pragma unroll
    for (int j = 0; j < 4; j++) pl[j] = low.p3(j);
    nw_depth3(pl, q, mix2, l, a, b, R);
    return R;
}
// blocks of tree 1 held in registers (the fused candidate kernel computes them once per record)
struct NwLowRegs {
    NwBlocks B;
    __device__ __forceinline__ uint32_t p1() const { return B.p1; }
    __device__ __forceinline__ uint32_t p2() const { return B.p2; }
    __device__ __forceinline__ uint32_t p3(int j) const { return B.p3[j]; }
};
#endif
