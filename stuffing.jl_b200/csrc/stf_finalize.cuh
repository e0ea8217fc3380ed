// stf_finalize.cuh — single-CTA fused kernels of the latency path (one scene, <= ~11k records / ~5.6k hits):
//   k_sort_records  : the spatial tree as sorted (cell, label) records          (qtrees.jl:399-552 containers)
//   k_finalize_hits : sort!(r) + unique! of the many-body result (qtree_functions.jl:252), its export, and the
//                     step-dependency links batchsteps needs (previous step of every movable object)
// Each replaces 4-8 launches of the throughput path; the counts stay on the device (no host round trip inside a
// fit round).  When a size does not fit, the kernels raise a status bit and the host re-runs the throughput path.
#pragma once
#include "stf_common.cuh"
#include "stf_blocksort.cuh"
#include "stf_clustersort.cuh"

#define STF_ST_RETRY_GROW 1   // an output buffer was too small: grow and re-run the round
#define STF_ST_RETRY_BIG 2    // too many records / hits for the single-CTA kernels: use the multi-launch path
#define FIN_IDX_BITS 14       // BS_CAP < 2^14

__global__ void __launch_bounds__(BS_THREADS) k_sort_records(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ tie_in,
                                                             const unsigned long long* __restrict__ n_dev, int64_t cap, int key_bits, int vbits,
                                                             const int32_t* __restrict__ val_map, uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                             unsigned long long* __restrict__ n_out, int* __restrict__ status) {
    pdl_sync();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BlockSortSmem& S = *reinterpret_cast<BlockSortSmem*>(smem_raw);
    int64_t n64 = (int64_t)*n_dev;
    if (n64 > cap || n64 > BS_CAP) { // (cap overflow cannot happen for records: the buffers hold 4 slots per object)
        if (threadIdx.x == 0) { atomicOr(status, STF_ST_RETRY_BIG); *n_out = 0; }
        return;
    }
    int n = (int)n64;
    int ch = bs_chunks(n);
    uint64_t x[BS_ITEMS];
#pragma unroll
    for (int c = 0; c < BS_ITEMS; c++) {
        int e = bs_index(n, c);
        x[c] = 0;
        if (c < ch && e < n) x[c] = (keys_in[e] << vbits) | (uint64_t)tie_in[e];
    }
    bs_sort(S, x, n, 0, key_bits + vbits); // key major, tie-break (label or position in `labels`) minor: a total order,
    uint64_t vmask = (1ull << vbits) - 1;  // so the unordered (atomic-append) input gives a deterministic result
    for (int e = threadIdx.x; e < n; e += BS_THREADS) {
        uint64_t p = S.buf[e];
        uint32_t v = (uint32_t)(p & vmask);
        keys_out[e] = p >> vbits;
        vals_out[e] = val_map ? (uint32_t)val_map[v] : v;
    }
    if (threadIdx.x == 0) *n_out = (unsigned long long)n;
}

// the same on a thread-block cluster (8 CTAs): CTA c sorts and writes the records [c*P, (c+1)*P)
__global__ void __cluster_dims__(CS_CTAS, 1, 1) __launch_bounds__(CS_THREADS) k_sort_records_cl(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ tie_in,
                                                             const unsigned long long* __restrict__ n_dev, int64_t cap, int key_bits, int vbits,
                                                             const int32_t* __restrict__ val_map, uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                             unsigned long long* __restrict__ n_out, int* __restrict__ status) {
    pdl_sync(); // programmatic dependent launch: scheduled while the previous kernel drains; nothing global before this
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ClusterSortSmem& S = *reinterpret_cast<ClusterSortSmem*>(smem_raw);
    const int c = (int)cg::this_cluster().block_rank();
    int64_t n64 = (int64_t)*n_dev;
    if (n64 > cap || n64 > CS_CAP) {
        if (c == 0 && threadIdx.x == 0) { atomicOr(status, STF_ST_RETRY_BIG); *n_out = 0; }
        return;
    }
    const int n = (int)n64, P = cs_slice(n), nl = min(max(n - c * P, 0), P);
    uint64_t x[CS_ITEMS];
#pragma unroll
    for (int j = 0; j < CS_ITEMS; j++) {
        int e = cs_local(j);
        x[j] = 0;
        if (e < nl) x[j] = (keys_in[c * P + e] << vbits) | (uint64_t)tie_in[c * P + e];
    }
    cs_sort(S, x, n, 0, key_bits + vbits);
    const uint64_t vmask = (1ull << vbits) - 1;
    for (int e = threadIdx.x; e < nl; e += CS_THREADS) {
        uint64_t p = S.buf[e];
        uint32_t v = (uint32_t)(p & vmask);
        keys_out[c * P + e] = p >> vbits;
        vals_out[c * P + e] = val_map ? (uint32_t)val_map[v] : v;
    }
    if (c == 0 && threadIdx.x == 0) *n_out = (unsigned long long)n;
}

struct FinalizeArgs {
    const Item16* hits; const unsigned long long* nh_dev; int64_t hits_cap;     // unordered hit list (device count)
    const unsigned long long* n_items_dev; int64_t item_cap;                    // overflow check of the candidate list
    Item16* sorted;                                                             // out: hits in canonical order
    int label_bits; int unique;
    int32_t* out5; unsigned long long* kept;                                    // optional export (5 ints per kept hit)
    int do_prep; const ObjMeta* om; uint32_t* prev; uint32_t* done;             // optional step-dependency links
    unsigned long long* ticket; unsigned long long* n_steps;
    int* status;
    const unsigned long long* stop; // optional: != 0 cancels the round as well (a multi-round call that has ended)
    int abort_on_status;   // != 0: a status bit that is already raised (by an earlier kernel of the round, by a peer GPU of an
                           // item-parallel group, or by an earlier round of a multi-round call) cancels the round: no step runs
};

__device__ __forceinline__ bool fin_lab_less(const Item16& x, const Item16& y) {
    int lx = item_l(x), ly = item_l(y);
    if (lx != ly) return lx < ly;
    if (x.a != y.a) return x.a < y.a;
    return x.b < y.b;
}

// step-dependency links of batchsteps (stf_step.cuh): endpoints (step p, side s) sorted by object, stable in list
// order; prev[2p+s] = previous step of that object + 1 (0 = none; the mask never moves: no dependency through it)
__device__ __forceinline__ void fin_step_links(BlockSortSmem& S, const Item16* __restrict__ list, int n, int lb, const ObjMeta* __restrict__ om,
                                               uint32_t* __restrict__ prev, uint32_t* __restrict__ done) {
    int n2 = 2 * n;
    uint64_t x[BS_ITEMS];
    {
        int ch = bs_chunks(n2);
#pragma unroll
        for (int c = 0; c < BS_ITEMS; c++) {
            int e = bs_index(n2, c);
            x[c] = 0;
            if (c < ch && e < n2) {
                Item16 h = list[e >> 1];
                uint32_t o = (uint32_t)((e & 1) ? item_i2(h) : h.i1) - 1u;
                uint64_t ok = om[o].is_mask ? (1ull << lb) : (uint64_t)o;
                x[c] = (ok << FIN_IDX_BITS) | (uint64_t)e;
            }
        }
    }
    __syncthreads();
    bs_sort(S, x, n2, FIN_IDX_BITS, FIN_IDX_BITS + lb + 1);
    for (int j = threadIdx.x; j < n2; j += BS_THREADS) {
        uint64_t cur = S.buf[j];
        uint64_t ok = cur >> FIN_IDX_BITS; uint32_t e = (uint32_t)(cur & ((1u << FIN_IDX_BITS) - 1));
        uint32_t pv = 0;
        if (ok != (1ull << lb) && j > 0) { uint64_t pr = S.buf[j - 1]; if ((pr >> FIN_IDX_BITS) == ok) pv = ((uint32_t)(pr & ((1u << FIN_IDX_BITS) - 1)) >> 1) + 1; }
        prev[e] = pv;
    }
    for (int p = threadIdx.x; p < n; p += BS_THREADS) done[p] = 0;
}

__global__ void __launch_bounds__(BS_THREADS) k_finalize_hits(FinalizeArgs A) {
    pdl_sync();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BlockSortSmem& S = *reinterpret_cast<BlockSortSmem*>(smem_raw);
    __shared__ uint32_t s_kept;
    int64_t nh64 = (int64_t)*A.nh_dev;
    bool grow = nh64 > A.hits_cap || (A.n_items_dev && (int64_t)*A.n_items_dev > A.item_cap);
    bool big = nh64 > (A.do_prep ? BS_CAP / 2 : BS_CAP);
    if (((A.abort_on_status && *A.status != 0) || (A.stop && *A.stop != 0ull)) && !grow && !big) { // cancelled from outside: keep the bits as they are
        if (threadIdx.x == 0) { if (A.n_steps) *A.n_steps = 0; if (A.ticket) *A.ticket = 0; if (A.kept) *A.kept = 0; }
        return;
    }
    if (grow || big) {
        if (threadIdx.x == 0) {
            atomicOr(A.status, grow ? STF_ST_RETRY_GROW : STF_ST_RETRY_BIG);
            if (A.n_steps) *A.n_steps = 0;
            if (A.ticket) *A.ticket = 0;
            if (A.kept) *A.kept = 0;
        }
        return;
    }
    int n = (int)nh64;
    const int lb = A.label_bits;
    // ---- sort by (i1, i2); payload = position in the unordered list
    uint64_t x[BS_ITEMS];
    {
        int ch = bs_chunks(n);
#pragma unroll
        for (int c = 0; c < BS_ITEMS; c++) {
            int e = bs_index(n, c);
            x[c] = 0;
            if (c < ch && e < n) {
                Item16 h = A.hits[e];
                x[c] = ((uint64_t)(uint32_t)h.i1 << (lb + FIN_IDX_BITS)) | ((uint64_t)(uint32_t)item_i2(h) << FIN_IDX_BITS) | (uint64_t)e;
            }
        }
    }
    bs_sort(S, x, n, FIN_IDX_BITS, FIN_IDX_BITS + 2 * lb);
    // ---- gather into canonical order
    for (int p = threadIdx.x; p < n; p += BS_THREADS) A.sorted[p] = A.hits[(uint32_t)(S.buf[p] & ((1u << FIN_IDX_BITS) - 1))];
    __syncthreads();
    // ---- hits of one ordered pair are adjacent: order each run by (l,a,b)  (sort! by ((i1,i2),(l,a,b)))
    for (int p = threadIdx.x; p < n; p += BS_THREADS) {
        uint64_t key = S.buf[p] >> FIN_IDX_BITS;
        bool head = p == 0 || (S.buf[p - 1] >> FIN_IDX_BITS) != key;
        if (!head) continue;
        int e = p + 1;
        while (e < n && (S.buf[e] >> FIN_IDX_BITS) == key) e++;
        for (int i = p + 1; i < e; i++) { // insertion sort of a (short) run
            Item16 v = A.sorted[i]; int y = i - 1;
            while (y >= p && fin_lab_less(v, A.sorted[y])) { A.sorted[y + 1] = A.sorted[y]; y--; }
            A.sorted[y + 1] = v;
        }
    }
    __syncthreads();
    // ---- export: unique!(minmax∘first) keeps the first entry of a run unless the swapped pair sorts earlier
    if (A.out5) {
        int per_t = (n + BS_THREADS - 1) / BS_THREADS, beg = threadIdx.x * per_t, end = min(n, beg + per_t);
        uint64_t lmask = (1ull << lb) - 1;
        auto keep_of = [&](int p) -> bool {
            if (!A.unique) return true;
            uint64_t key = S.buf[p] >> FIN_IDX_BITS;
            if (p > 0 && (S.buf[p - 1] >> FIN_IDX_BITS) == key) return false;
            uint64_t i1 = key >> lb, i2 = key & lmask;
            if (i1 > i2) {
                uint64_t sw = (i2 << lb) | i1;
                int lo = 0, hi = n;
                while (lo < hi) { int mid = (lo + hi) >> 1; if ((S.buf[mid] >> FIN_IDX_BITS) < sw) lo = mid + 1; else hi = mid; }
                if (lo < n && (S.buf[lo] >> FIN_IDX_BITS) == sw) return false;
            }
            return true;
        };
        uint32_t cnt = 0;
        for (int p = beg; p < end; p++) cnt += keep_of(p);
        uint32_t pos = bs_block_exclusive(cnt, S.wsum, &s_kept);
        for (int p = beg; p < end; p++) {
            if (!keep_of(p)) continue;
            Item16 h = A.sorted[p];
            int32_t* o = A.out5 + 5 * (size_t)pos++;
            o[0] = h.i1; o[1] = item_i2(h); o[2] = item_l(h); o[3] = h.a; o[4] = h.b;
        }
        if (threadIdx.x == 0 && A.kept) *A.kept = s_kept;
    }
    if (!A.do_prep) return;
    __syncthreads();
    fin_step_links(S, A.sorted, n, lb, A.om, A.prev, A.done);
    if (threadIdx.x == 0) { *A.ticket = 0; *A.n_steps = (unsigned long long)n; }
}

// links only, for a caller-ordered list (stf_batchsteps: the list order is the canonical step order); n <= BS_CAP/2
__global__ void __launch_bounds__(BS_THREADS) k_step_links(const Item16* __restrict__ list, int n, int lb, const ObjMeta* __restrict__ om,
                                                           uint32_t* __restrict__ prev, uint32_t* __restrict__ done, unsigned long long* __restrict__ ticket) {
    pdl_sync();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BlockSortSmem& S = *reinterpret_cast<BlockSortSmem*>(smem_raw);
    fin_step_links(S, list, n, lb, om, prev, done);
    if (threadIdx.x == 0) *ticket = 0;
}

// ---- the fit-round finalize on a thread-block cluster (no export): sort!(hits) -> canonical order -> tie order ->
// step-dependency links.  CTA c owns slice c of every sorted sequence; neighbours across a slice boundary are read from
// global memory (sorted keys) or from the previous CTA's shared memory (endpoint records).
__global__ void __cluster_dims__(CS_CTAS, 1, 1) __launch_bounds__(CS_THREADS) k_finalize_prep_cl(FinalizeArgs A, uint64_t* __restrict__ skey) {
    pdl_sync(); // programmatic dependent launch: scheduled while the previous kernel drains; nothing global before this
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ClusterSortSmem& S = *reinterpret_cast<ClusterSortSmem*>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    int64_t nh64 = (int64_t)*A.nh_dev;
    bool grow = nh64 > A.hits_cap || (A.n_items_dev && (int64_t)*A.n_items_dev > A.item_cap);
    bool big = nh64 > CS_CAP / 2;
    if (((A.abort_on_status && *A.status != 0) || (A.stop && *A.stop != 0ull)) && !grow && !big) { // cancelled from outside: keep the bits as they are
        if (c == 0 && threadIdx.x == 0) { *A.n_steps = 0; *A.ticket = 0; }
        return;
    }
    if (grow || big) {
        if (c == 0 && threadIdx.x == 0) { atomicOr(A.status, grow ? STF_ST_RETRY_GROW : STF_ST_RETRY_BIG); *A.n_steps = 0; *A.ticket = 0; }
        return;
    }
    const int n = (int)nh64, lb = A.label_bits;
    uint64_t x[CS_ITEMS];
    { // ---- sort by (i1, i2); payload = position in the unordered list
        const int P = cs_slice(n), nl = min(max(n - c * P, 0), P);
#pragma unroll
        for (int j = 0; j < CS_ITEMS; j++) {
            int e = cs_local(j);
            x[j] = 0;
            if (e < nl) { Item16 h = A.hits[c * P + e]; x[j] = ((uint64_t)(uint32_t)h.i1 << (lb + FIN_IDX_BITS)) | ((uint64_t)(uint32_t)item_i2(h) << FIN_IDX_BITS) | (uint64_t)(c * P + e); }
        }
        cs_sort(S, x, n, FIN_IDX_BITS, FIN_IDX_BITS + 2 * lb);
        for (int e = threadIdx.x; e < nl; e += CS_THREADS) { // gather into canonical order; keys to global for the neighbours
            uint64_t p = S.buf[e];
            A.sorted[c * P + e] = A.hits[(uint32_t)(p & ((1u << FIN_IDX_BITS) - 1))];
            skey[c * P + e] = p >> FIN_IDX_BITS;
        }
        cluster.sync();
        for (int e = threadIdx.x; e < nl; e += CS_THREADS) { // order each run of one ordered pair by (l,a,b)
            int p = c * P + e;
            uint64_t key = skey[p];
            if (p > 0 && skey[p - 1] == key) continue;
            int q = p + 1;
            while (q < n && skey[q] == key) q++;
            for (int i = p + 1; i < q; i++) {
                Item16 v = A.sorted[i]; int y = i - 1;
                while (y >= p && fin_lab_less(v, A.sorted[y])) { A.sorted[y + 1] = A.sorted[y]; y--; }
                A.sorted[y + 1] = v;
            }
        }
        cluster.sync();
    }
    { // ---- endpoints (step p, side s) sorted by object, stable in list order -> prev links
        const int n2 = 2 * n, P = cs_slice(n2), nl = min(max(n2 - c * P, 0), P);
#pragma unroll
        for (int j = 0; j < CS_ITEMS; j++) {
            int e = cs_local(j);
            x[j] = 0;
            if (e < nl) {
                int g = c * P + e;
                Item16 h = A.sorted[g >> 1];
                uint32_t o = (uint32_t)((g & 1) ? item_i2(h) : h.i1) - 1u;
                uint64_t ok = A.om[o].is_mask ? (1ull << lb) : (uint64_t)o;
                x[j] = (ok << FIN_IDX_BITS) | (uint64_t)g;
            }
        }
        cs_sort(S, x, n2, FIN_IDX_BITS, FIN_IDX_BITS + lb + 1);
        for (int e = threadIdx.x; e < nl; e += CS_THREADS) {
            uint64_t cur = S.buf[e];
            uint64_t ok = cur >> FIN_IDX_BITS; uint32_t g = (uint32_t)(cur & ((1u << FIN_IDX_BITS) - 1));
            uint32_t pv = 0;
            if (ok != (1ull << lb) && (e > 0 || c > 0)) {
                uint64_t pr = e > 0 ? S.buf[e - 1] : cluster.map_shared_rank(&S, c - 1)->buf[P - 1]; // the slice before is full
                if ((pr >> FIN_IDX_BITS) == ok) pv = ((uint32_t)(pr & ((1u << FIN_IDX_BITS) - 1)) >> 1) + 1;
            }
            A.prev[g] = pv;
        }
        for (int p = c * CS_THREADS + threadIdx.x; p < n; p += CS_CTAS * CS_THREADS) A.done[p] = 0;
        if (c == 0 && threadIdx.x == 0) { *A.ticket = 0; *A.n_steps = (unsigned long long)n; }
        cluster.sync(); // nobody leaves while a neighbour may still read its shared memory
    }
}
