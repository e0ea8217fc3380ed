// stf_pyramid.cuh — subsystem (1): pack + pyramid build / rebuild kernels.
// Replaces PaddedMat/ShiftedQTree construction and buildqtree! (qtrees.jl:117-128,176-237) and the
// rebuild half of shift!/setshift! (qtrees.jl:238-287).
#pragma once
#include "stf_common.cuh"

// ---- level-1 pack: u8 codes (column-major payload, pitch kh) -> 2-bit words -------------------
// one thread per output word; `woff1[o]` = first level-1 word of object o in a dense numbering
// (prefix sum over objects), `poff[o]` = byte offset of its payload in `payload`.
__global__ void k_pack_level1(const uint8_t* __restrict__ payload, const uint64_t* __restrict__ poff,
                              const uint64_t* __restrict__ woff1, int n, const LevelMeta* __restrict__ lm, int L,
                              uint32_t* __restrict__ words, uint64_t total_words, int* __restrict__ err) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_words) return;
    int lo = 0, hi = n; // largest o with woff1[o] <= i
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (woff1[mid] <= i) lo = mid; else hi = mid; }
    int o = lo;
    LevelMeta m = lm[(size_t)o * L];
    int kh = lm_kh(m), wpc = wpc_of(kh);
    uint32_t local = (uint32_t)(i - woff1[o]);
    int c = local / wpc, w = local - c * wpc;
    const uint8_t* src = payload + poff[o] + (size_t)c * kh + 16 * w;
    int nv = min(16, kh - 16 * w);
    uint32_t d = lm_dflt(m), out = 0; int bad = 0;
    if (nv == 16) { // 16 payload bytes at any alignment: 5 aligned words + funnel shifts (buffer has slack)
        const uint32_t* ap = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
        int sh = (int)(reinterpret_cast<uintptr_t>(src) & 3) * 8;
        uint32_t t[5];
#pragma unroll
        for (int k = 0; k < 5; k++) t[k] = ap[k];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t x = __funnelshift_r(t[k], t[k + 1], sh);
            bad |= ((x & 0xFCFCFCFCu) != 0) | (((x | (x >> 1)) & 0x01010101u) != 0x01010101u);
            uint32_t y = (x & 3u) | ((x >> 6) & 0xCu) | ((x >> 12) & 0x30u) | ((x >> 18) & 0xC0u);
            out |= y << (8 * k);
        }
    } else {
        for (int j = 0; j < 16; j++) {
            uint32_t code = j < nv ? src[j] : d;
            bad |= (code < 1u || code > 3u);
            out |= (code & 3u) << (2 * j);
        }
    }
    if (bad) atomicOr(err, 1);
    words[m.off + local] = out;
}

// ---- warp-per-object build: worklist entry = (object, from_level): levels from+1..L ------------
__global__ void k_build_objects(uint32_t* __restrict__ words, LevelMeta* __restrict__ lm, int L,
                                const int2* __restrict__ work, int nwork, const ObjMeta* __restrict__ om, int skip_big) {
    __shared__ uint32_t sbuf[8][2 * RB_HALF]; // blockDim.x == 256
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < nwork; i += nwarps) {
        int2 w = work[i];
        if (skip_big && om[w.x].big) continue;
        if (w.y < L) rebuild_levels_warp<false>(words, lm, w.x, L, w.y, lane, sbuf[threadIdx.x >> 5]);
        __syncwarp();
    }
}

// ---- wide build of ONE level of ONE (big) object: grid-stride over parent words ----------------
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
}
// the shift of level l must be set before the NEXT level reads it; done in its own tiny launch so
// that no block of k_build_level_wide races with it
__global__ void k_set_parent_shift(LevelMeta* __restrict__ lm, int L, int o, int l) {
    LevelMeta ch = lm[(size_t)o * L + (l - 2)];
    lm[(size_t)o * L + (l - 1)].rs = ch.rs / 2;
    lm[(size_t)o * L + (l - 1)].cs = ch.cs / 2;
}

// ---- shift!/setshift! (qtrees.jl:267-287): levels <= level move by s*2^(level-i); build list -----
__global__ void k_apply_shifts(LevelMeta* __restrict__ lm, int L, int n, const int32_t* __restrict__ labels, int level,
                               const int32_t* __restrict__ rs, const int32_t* __restrict__ cs, int add, int2* __restrict__ work) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int o = labels ? labels[i] - 1 : i;
    int s1 = rs[i], s2 = cs[i];
    for (int l = level; l >= 1; l--) {
        LevelMeta* m = lm + (size_t)o * L + (l - 1);
        if (add) { m->rs += s1; m->cs += s2; } else { m->rs = s1; m->cs = s2; }
        s1 *= 2; s2 *= 2;
    }
    work[i] = make_int2(o, level);
}
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
