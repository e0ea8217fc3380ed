// stf_collide.cuh — subsystems (2) broad phase and (3) narrow phase.
//  (2) locate!/positionlower (qtree_functions.jl:141-201), the spatial trees as sorted (cell,label)
//      records, candidate generation of totalcollisions_spacial (:230-249), collisions_boundsfilter
//      (:203-223) and partialcollisions (:285-326).
//  (3) _collision_randbfs / collision / collision_dfs (:2-51) as a level-synchronous frontier.
#pragma once
#include <cooperative_groups.h>
#include "stf_common.cuh"
#include "stf_narrow.cuh"
namespace cg = cooperative_groups;

// =============================================================================================
// (2) broad phase
// =============================================================================================

// FOUR lanes per label (<= 4 cells): the four top cells, then the four children of every positionlower step, are read by
// the four lanes at once and decided by one ballot, so a descent costs one round trip per level instead of up to five
// (the thread-per-label version spent 19 us of a 300 us round in this dependent chain).  The warp runs one converged loop.
// Slot mode (n_append == NULL): written to slots 4*slot..4*slot+3 (unused slots = INVALID key).
//   linked != 0: LinkedSpacialQTree semantics — slot = object, cells outside the root (L,1,1) dropped (qtrees.jl:506).
//   linked == 0: HashSpacialQTree semantics — slot = position in `labels`.
// Append mode (n_append != NULL, hash semantics): the valid records are appended compactly (one atomic per warp) as
//   (key, tie) with tie = position in `labels` (or the label itself when labels == NULL); a sort by (key, tie)
//   reproduces the insertion order of the reference's Dict of Vectors (qtrees.jl:411-415) whatever the append order.
#define LOC_THREADS 128
// device-side control of a round inside a multi-round call (stf_fit_rounds, stf_fit_epoch_e): `stop` != 0 cancels the round
// (no records are produced, so every later kernel finds empty lists); `nl_dev` = the number of labels when the label list
// was produced on the device (the labels of the previous round's colist)
struct RoundCtl { const unsigned long long* stop; const int* nl_dev; };
__global__ void __launch_bounds__(LOC_THREADS) k_locate(const uint32_t* __restrict__ words, const LevelMeta* __restrict__ lm, const ObjMeta* __restrict__ om, int L,
                         int nl, const int32_t* __restrict__ labels, int linked, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                         int* __restrict__ err, unsigned long long* __restrict__ n_append, RoundCtl ctl) {
    pdl_sync(); // programmatic dependent launch: scheduled while the previous kernel drains; nothing global before this
    if (ctl.stop && *ctl.stop) return; // a cancelled round of a multi-round call: no records, so every later kernel finds empty lists
    if (ctl.nl_dev) nl = min(nl, *ctl.nl_dev);
    const unsigned FULLM = 0xffffffffu;
    const int lane = threadIdx.x & 31, gl = lane & 3, g0 = lane & ~3;
    const int j = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 2);
    const bool have = j < nl;
    int cnt = 0; uint64_t ck0 = STF_KEY_INVALID, ck1 = STF_KEY_INVALID, ck2 = STF_KEY_INVALID, ck3 = STF_KEY_INVALID;
    int o = 0, slot = 0, scene = 0;
    const LevelMeta* olm = lm;
    unsigned topmask = 0; LevelMeta top = {};
    if (have) {
        o = labels ? labels[j] - 1 : j;
        slot = linked ? o : j;
        olm = lm + (size_t)o * L;
        top = olm[L - 1]; // sizelevel(qt) == length(qt)  qtrees.jl:191-193
        scene = om[o].scene;
    }
    { // the four top cells (1,1),(1,2),(2,1),(2,2) of the top kernel, lane k reads cell k  qtree_functions.jl:167-170
        bool ne = have && node_code<false>(words, top, top.rs + 1 + (gl >> 1), top.cs + 1 + (gl & 1)) != CODE_EMPTY;
        topmask = (__ballot_sync(FULLM, ne) >> g0) & 0xFu;
    }
    bool indesc = false; int l = 0, a = 0, b = 0;
    LevelMeta nextm = {}; // meta of the level the NEXT iteration reads: loaded one iteration ahead, off the node-read chain
    if (have && L > 2) nextm = olm[L - 2];
    for (;;) {
        if (!indesc && topmask) { // next non-empty top cell, in order
            int k = __ffs(topmask) - 1; topmask &= topmask - 1;
            if (l != 0 && L > 2) nextm = olm[L - 2]; // a second top cell restarts from the top (rare)
            l = L; a = top.rs + 1 + (k >> 1); b = top.cs + 1 + (k & 1); indesc = true;
        }
        if (!__any_sync(FULLM, indesc)) break;
        const bool go = indesc && l > 2; // positionlower  :141-158: one level per iteration, lane i tests child i
        bool ne = false;
        if (go) {
            LevelMeta cm = nextm;             // = olm[l - 2]
            if (l > 3) nextm = olm[l - 3];    // the level after this one, in flight together with the node read
            ne = node_code<false>(words, cm, 2 * a - (gl & 1), 2 * b - (gl >> 1)) != CODE_EMPTY;
        }
        const unsigned nem = (__ballot_sync(FULLM, ne) >> g0) & 0xFu;
        if (indesc) {
            bool fin = !go;
            if (go) {
                if (__popc(nem) == 1) { int i = __ffs(nem) - 1; a = 2 * a - (i & 1); b = 2 * b - (i >> 1); l -= 1; } // the only non-empty child
                else fin = true; // two non-empty children: stay; none: inconsistent pyramid (the oracle stops the same way)
            }
            if (fin) {
                bool keep = true;
                if (linked) { int r = 1 << (L - l); keep = (a >= 1 && a <= r && b >= 1 && b <= r); }
                if (a <= -STF_CELL_BIAS || a >= STF_CELL_BIAS || b <= -STF_CELL_BIAS || b >= STF_CELL_BIAS) { if (gl == 0) atomicOr(err, 2); keep = false; }
                if (keep) {
                    uint64_t key = cell_key(scene, l, a, b);
                    if (cnt == 0) ck0 = key; else if (cnt == 1) ck1 = key; else if (cnt == 2) ck2 = key; else ck3 = key;
                    cnt++;
                }
                indesc = false;
            }
        }
    }
    const uint64_t mine = gl == 0 ? ck0 : gl == 1 ? ck1 : gl == 2 ? ck2 : ck3; // lane k of the group holds record k
    if (n_append) {
        int c = gl == 0 ? cnt : 0, incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULLM, incl, d); if (lane >= d) incl += t; }
        int total = __shfl_sync(FULLM, incl, 31);
        unsigned long long base = 0;
        if (lane == 31 && total) base = atomicAdd(n_append, (unsigned long long)total);
        base = __shfl_sync(FULLM, base, 31);
        int excl = __shfl_sync(FULLM, incl - c, g0); // records of the groups before mine in this warp
        uint32_t tie = labels ? (uint32_t)j : (uint32_t)(o + 1);
        if (gl < cnt) { keys[base + excl + gl] = mine; vals[base + excl + gl] = tie; }
        return;
    }
    if (!have) return;
    keys[4 * (size_t)slot + gl] = gl < cnt ? mine : STF_KEY_INVALID; vals[4 * (size_t)slot + gl] = (uint32_t)(o + 1);
}

// Throughput version, one THREAD per label (<= 4 cells): batches of scenes and very large scenes, where every SM is
// busy anyway and the four-lane version would only multiply the instruction count.
// Slot mode (n_append == NULL): written to slots 4*slot..4*slot+3 (unused slots = INVALID key).
//   linked != 0: LinkedSpacialQTree semantics — slot = object, cells outside the root (L,1,1) dropped (qtrees.jl:506).
//   linked == 0: HashSpacialQTree semantics — slot = position in `labels`.
// Append mode (n_append != NULL, hash semantics): the valid records are appended compactly (one atomic per warp) as
//   (key, tie) with tie = position in `labels` (or the label itself when labels == NULL); a sort by (key, tie)
//   reproduces the insertion order of the reference's Dict of Vectors (qtrees.jl:411-415) whatever the append order.
__global__ void k_locate_wide(const uint32_t* __restrict__ words, const LevelMeta* __restrict__ lm, const ObjMeta* __restrict__ om, int L,
                         int nl, const int32_t* __restrict__ labels, int linked, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                         int* __restrict__ err, unsigned long long* __restrict__ n_append, RoundCtl ctl) {
    pdl_sync(); // programmatic dependent launch: scheduled while the previous kernel drains; nothing global before this
    if (ctl.stop && *ctl.stop) return;
    if (ctl.nl_dev) nl = min(nl, *ctl.nl_dev);
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0; uint64_t ck[4]; int o = 0, slot = 0;
    if (j < nl) {
        o = labels ? labels[j] - 1 : j;
        slot = linked ? o : j;
        const LevelMeta* olm = lm + (size_t)o * L;
        LevelMeta top = olm[L - 1]; // sizelevel(qt) == length(qt)  qtrees.jl:191-193
        int scene = om[o].scene;
        for (int k = 0; k < 4; k++) { // (1,1),(1,2),(2,1),(2,2)  qtree_functions.jl:167-170
            int l = L, a = top.rs + 1 + (k >> 1), b = top.cs + 1 + (k & 1);
            if (node_code<false>(words, top, a, b) == CODE_EMPTY) continue;
            while (l > 2) { // positionlower  :141-158
                LevelMeta cm = olm[l - 2];
                int ca = a, cb = b, flag = 0, stop = 0;
                for (int i = 0; i < 4; i++) {
                    int xa = 2 * a - (i & 1), xb = 2 * b - (i >> 1);
                    if (node_code<false>(words, cm, xa, xb) != CODE_EMPTY) {
                        if (flag) { stop = 1; break; }
                        flag = 1; ca = xa; cb = xb;
                    }
                }
                if (stop || !flag) break; // !flag: inconsistent pyramid; the reference would not terminate (oracle: same stop)
                l = l - 1; a = ca; b = cb;
            }
            bool keep = true;
            if (linked) { int r = 1 << (L - l); keep = (a >= 1 && a <= r && b >= 1 && b <= r); }
            if (a <= -STF_CELL_BIAS || a >= STF_CELL_BIAS || b <= -STF_CELL_BIAS || b >= STF_CELL_BIAS) { atomicOr(err, 2); keep = false; }
            if (keep) ck[cnt++] = cell_key(scene, l, a, b);
        }
    }
    if (n_append) {
        const unsigned FULLM = 0xffffffffu;
        int lane = threadIdx.x & 31, incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULLM, incl, d); if (lane >= d) incl += t; }
        int total = __shfl_sync(FULLM, incl, 31);
        unsigned long long base = 0;
        if (lane == 31 && total) base = atomicAdd(n_append, (unsigned long long)total);
        base = __shfl_sync(FULLM, base, 31) + (unsigned long long)(incl - cnt);
        uint32_t tie = labels ? (uint32_t)j : (uint32_t)(o + 1);
        for (int k = 0; k < cnt; k++) { keys[base + k] = ck[k]; vals[base + k] = tie; }
        return;
    }
    if (j >= nl) return;
    for (int k = 0; k < 4; k++) { keys[4 * (size_t)slot + k] = k < cnt ? ck[k] : STF_KEY_INVALID; vals[4 * (size_t)slot + k] = (uint32_t)(o + 1); }
}
// LinkedSpacialQTree table (4 slots per object) -> compact (key, label) records, unordered
__global__ void k_compact_records(const uint64_t* __restrict__ tk, const uint32_t* __restrict__ tv, int64_t nslots,
                                  uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, unsigned long long* __restrict__ n_append,
                                  RoundCtl ctl) {
    pdl_sync();
    if (ctl.stop && *ctl.stop) return;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t k = i < nslots ? tk[i] : STF_KEY_INVALID;
    bool v = k != STF_KEY_INVALID;
    unsigned m = __ballot_sync(0xffffffffu, v);
    int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0 && m) base = atomicAdd(n_append, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0) + __popc(m & ((1u << lane) - 1));
    if (v) { keys[base] = k; vals[base] = tv[i]; }
}

__global__ void k_fill_invalid(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = STF_KEY_INVALID; vals[i] = (uint32_t)(i / 4 + 1); }
}
__global__ void k_clear_slots(uint64_t* __restrict__ keys, int nl, const int32_t* __restrict__ labels) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nl) { size_t o = (size_t)(labels[j] - 1); for (int k = 0; k < 4; k++) keys[4 * o + k] = STF_KEY_INVALID; }
}

__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t* __restrict__ keys, int64_t n, uint64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// (scene, level) index of the sorted records: idx[sl] = first record whose (scene << 5 | level) >= sl, idx[nsl] = n.
// Batches of scenes search an ancestor cell inside its own (scene, level) slice (a few records) instead of the whole array.
__global__ void k_level_index(const uint64_t* __restrict__ keys, const unsigned long long* __restrict__ n_dev, int64_t cap, int32_t* __restrict__ idx, int nsl) {
    pdl_sync();
    int64_t n = min((int64_t)*n_dev, cap);
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    int64_t prev = r == 0 ? -1 : (int64_t)(keys[r - 1] >> (2 * STF_CELL_W));
    int64_t cur = r == n ? nsl : (int64_t)(keys[r] >> (2 * STF_CELL_W));
    if (cur > nsl) cur = nsl;
    for (int64_t x = prev + 1; x <= cur; x++) idx[x] = (int32_t)r;
}

// a candidate pair the word-parallel part could not decide (or that overflowed the small frontier): the item, its
// frontier `front` (Morton order of the child-number paths) at `depth` levels below the start node, and — for explicit
// item lists — the index its result is stored under
struct __align__(16) SurvRec { Item16 item; uint64_t front; uint32_t depth; uint32_t it; };

struct EmitOut {
    Item16* items; int64_t item_cap;      // materialised candidate list (only when FUSED == false)
    Item16* direct; int64_t direct_cap;   // direct hits AND (fused) narrow-phase hits go straight into the hit list ...
    unsigned long long* counts;           // [0] items, [1] direct hits, [3] node tests (statistics)
    unsigned long long* nhits;            // ... whose fill counter this is
    SurvRec* surv; int64_t surv_cap; unsigned long long* nsurv; // fused: pairs still undecided after depth 3
    unsigned long long* cursor;           // work cursor of the frontier kernel (zeroed here)
};
__device__ __forceinline__ void emit_push(Item16* buf, int64_t cap, unsigned long long* counter, const Item16& it) {
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned long long base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(counter, (unsigned long long)g.size());
    base = g.shfl(base, 0);
    unsigned long long slot = base + g.thread_rank();
    if ((int64_t)slot < cap) buf[slot] = it;
}

__device__ __forceinline__ int64_t upper_bound_u64(const uint64_t* __restrict__ keys, int64_t lo, int64_t hi, uint64_t key) {
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (keys[mid] <= key) lo = mid + 1; else hi = mid; }
    return lo;
}

// candidate generation, W lanes (a warp, or a half warp for batches of small scenes) per sorted record r (low label in
// cell (l,a,b)).
//  lane 0      : the later records of r's own cell            (same-cell pairs, qtree_functions.jl:233-239)
//  lane k >= 1 : the records of the k-th ancestor cell, found by binary search (:241-248)
// A scan turns the per-lane run lengths into one candidate range; the lanes then walk that range W candidates at a time
// (balanced even when an ancestor holds hundreds of labels), apply collisions_boundsfilter (:203-223) and reserve output
// slots with one atomic per warp iteration.  W = 16 when records have few candidates (a scene of a batch has ~10 per
// record and at most 16 levels): two records per warp keep twice as many lanes busy in the searches and in the filter.
// moved_pos != NULL selects partialcollisions semantics (:285-326): a pair is generated iff at least one
// side is in `labels` (moved_pos[o] >= 0); two moved labels of one cell are oriented (later, earlier).
// FUSED: the narrow phase of every generated pair runs right here, in the lane that generated it (nw_collide3: the
// first three levels below the cell as bitwise plane tests on whole node blocks): the candidate list is never written
// or re-read, the blocks of the LOW tree (the same for every candidate of the record) are fetched once per record by 14
// lanes (2 + 4 + 8 columns) and handed round by shuffles; decided pairs end here (hits appended to the hit list), the few
// that are still undecided after depth 3 go to the frontier kernel with their frontier mask.
#define EM_THREADS 256
template <int W, bool FUSED>
__global__ void __launch_bounds__(EM_THREADS) k_emit_items(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                           const unsigned long long* __restrict__ nrec_dev, int64_t nrec_cap, int L,
                                                           const uint32_t* __restrict__ words, const LevelMeta* __restrict__ lm,
                                                           const int32_t* __restrict__ moved_pos, EmitOut out, int shard_rank, int shard_n,
                                                           const int32_t* __restrict__ lvl_index) {
    pdl_sync(); // programmatic dependent launch: scheduled while the previous kernel drains; nothing global before this
    constexpr int RPW = 32 / W; // records per warp
    __shared__ int s_start[EM_THREADS / 32][RPW][16];
    __shared__ int s_pref[EM_THREADS / 32][RPW][17];
    const unsigned FULLM = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, sub = lane / W, sl_ = lane % W;
    int64_t nrec = min((int64_t)*nrec_dev, nrec_cap); // valid records (sorted, compacted)
    int64_t nwarps = (int64_t)gridDim.x * (EM_THREADS / 32);
    unsigned long long visited = 0;
    if (FUSED && blockIdx.x == 0 && threadIdx.x == 0) *out.cursor = 0ull;
    // item-parallel sharding: rank k of shard_n generates the candidates of the records k, k+shard_n, ...
    for (int64_t idx0 = ((int64_t)blockIdx.x * (EM_THREADS / 32) + wib) * RPW; idx0 * shard_n + shard_rank < nrec; idx0 += nwarps * RPW) {
        const int64_t r = (idx0 + sub) * shard_n + shard_rank;
        const bool rec = r < nrec; // (the second half of the last warp may have no record)
        uint64_t key = rec ? keys[r] : 0ull;
        int low = rec ? (int)vals[r] : 1;
        int l, a, b; cell_unkey(key, &l, &a, &b);
        int mp_low = moved_pos ? moved_pos[low - 1] : 0;
        // ---- per-lane candidate run
        int start = 0, cnt = 0;
        const int64_t sl = (int64_t)(key >> (2 * STF_CELL_W)); // scene << 5 | level
        // every lane looks for the run of ONE key inside [lo, hi): lane 0 its own key after r, lane j the ancestor cell j
        // levels up, inside that (scene, level) slice when the index is there.  One code path for all lanes (no serialised
        // lane-0 branch); the run end is found by galloping (runs are one or two records long).
        bool look = rec && (sl_ == 0 || (sl_ < L - 1 && l + sl_ <= L));
        uint64_t akey = key;
        int64_t lo = r + 1, hi = (lvl_index && rec) ? (int64_t)lvl_index[sl + 1] : nrec;
        if (look && sl_ > 0) {
            int pa = a, pb = b;
            for (int i = 0; i < sl_; i++) { pa = (pa + 1) / 2; pb = (pb + 1) / 2; } // parent(): ÷ truncates toward zero
            uint64_t scene_bits = key >> (5 + 2 * STF_CELL_W);
            akey = (scene_bits << (5 + 2 * STF_CELL_W)) | ((uint64_t)(uint32_t)(l + sl_) << (2 * STF_CELL_W)) |
                   ((uint64_t)(uint32_t)(pa + STF_CELL_BIAS) << STF_CELL_W) | (uint64_t)(uint32_t)(pb + STF_CELL_BIAS);
            lo = 0; hi = nrec;
            if (lvl_index) { lo = lvl_index[sl + sl_]; hi = lvl_index[sl + sl_ + 1]; } // the ancestor's own (scene, level) slice
            int64_t x = lo, y = hi;
            while (x < y) { int64_t mid = (x + y) >> 1; if (keys[mid] < akey) x = mid + 1; else y = mid; }
            lo = x;
        }
        if (look && lo < hi && keys[lo] == akey) { // run [lo, e): gallop, then bisect the last stride
            int64_t step = 1;
            while (lo + step < hi && keys[lo + step] == akey) step <<= 1;
            int64_t x = lo + (step >> 1) + 1, y = min(lo + step, hi); // keys[lo + step/2] == akey, keys[y] != akey (or y == hi)
            while (x < y) { int64_t mid = (x + y) >> 1; if (keys[mid] == akey) x = mid + 1; else y = mid; }
            start = (int)lo; cnt = (int)(x - lo);
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < W; o <<= 1) { int t = __shfl_up_sync(FULLM, incl, o, W); if (sl_ >= o) incl += t; }
        const int total = __shfl_sync(FULLM, incl, W - 1, W);
        int tmax = total;
        if (W == 16) tmax = max(tmax, __shfl_xor_sync(FULLM, total, 16));
        if (tmax == 0) continue;
        if (sl_ < 16) { s_start[wib][sub][sl_] = start; s_pref[wib][sub][sl_] = incl - cnt; }
        if (sl_ == 0) s_pref[wib][sub][16] = total;
        __syncwarp();
        LevelMeta lowm = {};
        if (rec) lowm = lm[(size_t)(low - 1) * L + (l - 1)];
        int lowcode = -1;
        NwLowRegs lowb = {};
        if (FUSED) { // the low tree's blocks of depth 1..3 below the cell: lane t < 14 of the record's group fetches one column
            const int d = sl_ < 2 ? 1 : (sl_ < 6 ? 2 : 3), j = sl_ < 2 ? sl_ : (sl_ < 6 ? sl_ - 2 : sl_ - 6);
            uint32_t fv = 0;
            if (rec && sl_ < 14 && l - d >= 1) {
                const LevelMeta m = lm[(size_t)(low - 1) * L + (l - d - 1)];
                const int n = 1 << d;
                fv = fields32_at<false>(words, m, n * b - (n - 1) - m.cs - 1 + j, n * a - (n - 1) - m.rs - 1, n);
            }
            const int g0 = sub * W;
#define STF_F(k) __shfl_sync(FULLM, fv, g0 + (k))
            lowb.B.p1 = (STF_F(0) & 0xFu) | ((STF_F(1) & 0xFu) << 4);
            lowb.B.p2 = (STF_F(2) & 0xFFu) | ((STF_F(3) & 0xFFu) << 8) | ((STF_F(4) & 0xFFu) << 16) | ((STF_F(5) & 0xFFu) << 24);
#pragma unroll
            for (int q = 0; q < 4; q++) lowb.B.p3[q] = (STF_F(6 + 2 * q) & 0xFFFFu) | (STF_F(7 + 2 * q) << 16);
#undef STF_F
        }
        const int* pref = s_pref[wib][sub]; const int* sstart = s_start[wib][sub];
        for (int c0 = 0; c0 < tmax; c0 += W) {
            int c = c0 + sl_;
            bool valid = c < total;
            bool is_item = false, is_direct = false, want_code = false;
            int i1 = low, i2 = 0;
            if (valid) {
                // the run candidate c belongs to: the LAST k with pref[k] <= c (runs are in lane order, empty runs share a
                // prefix with their successor, pref[0] = 0): four bisection steps over the 16 prefixes
                int k = pref[8] <= c ? 8 : 0;
                k += pref[k + 4] <= c ? 4 : 0;
                k += pref[k + 2] <= c ? 2 : 0;
                k += pref[k + 1] <= c ? 1 : 0;
                int other = (int)vals[sstart[k] + (c - pref[k])];
                if (k == 0) {
                    is_item = true; i2 = other;
                    if (moved_pos) {
                        int mp_o = moved_pos[other - 1];
                        if (mp_low < 0 && mp_o < 0) is_item = false;
                        else if (mp_low < 0 || (mp_o >= 0 && mp_o > mp_low)) { i1 = other; i2 = low; }
                    }
                } else if (!(moved_pos && mp_low < 0 && moved_pos[other - 1] < 0)) {
                    LevelMeta hm = lm[(size_t)(other - 1) * L + (l - 1)];
                    int rr = a - hm.rs - 1, cc = b - hm.cs - 1;
                    i2 = other;
                    if ((unsigned)rr < (unsigned)lm_kh(hm) && (unsigned)cc < (unsigned)lm_kw(hm)) is_item = true;
                    else if (lm_dflt(hm) == CODE_FULL) want_code = true;
                }
            }
            if (__any_sync(FULLM, want_code)) {
                if (lowcode < 0 && rec) lowcode = (int)node_code<false>(words, lowm, a, b);
                is_direct = want_code && lowcode != (int)CODE_EMPTY;
            }
            unsigned im = __ballot_sync(FULLM, is_item), dm = __ballot_sync(FULLM, is_direct);
            unsigned long long ibase = 0, dbase = 0;
            if (lane == 0) {
                if (im) ibase = atomicAdd(out.counts, (unsigned long long)__popc(im));
                if (dm) { dbase = atomicAdd(out.nhits, (unsigned long long)__popc(dm)); atomicAdd(out.counts + 1, (unsigned long long)__popc(dm)); }
            }
            ibase = __shfl_sync(FULLM, ibase, 0); dbase = __shfl_sync(FULLM, dbase, 0);
            unsigned lt = (1u << lane) - 1;
            if (!FUSED) { if (is_item) { int64_t slot = (int64_t)ibase + __popc(im & lt); if (slot < out.item_cap) out.items[slot] = make_item(i1, i2, l, a, b); } }
            else if (is_item) { // narrow phase in place: `low` against the other tree of the pair (the AND is symmetric)
                const int other = i1 == low ? i2 : i1;
                const NwResult R = nw_collide3<false>(words, lowb, lm + (size_t)(other - 1) * L, l, a, b);
                visited += (unsigned long long)R.visited;
                if (R.undecided) {
                    unsigned long long s_ = atomicAdd(out.nsurv, 1ull);
                    if ((int64_t)s_ < out.surv_cap) { SurvRec sr; sr.item = make_item(i1, i2, l, a, b); sr.front = R.front; sr.depth = 3u; sr.it = 0u; out.surv[s_] = sr; }
                } else if (R.l >= 0) { // cp[1] >= 0  qtree_functions.jl:104-106
                    unsigned long long s_ = atomicAdd(out.nhits, 1ull);
                    if ((int64_t)s_ < out.direct_cap) out.direct[s_] = make_item(i1, i2, R.l, R.a, R.b);
                }
            }
            if (is_direct) { int64_t slot = (int64_t)dbase + __popc(dm & lt); if (slot < out.direct_cap) out.direct[slot] = make_item(low, i2, l, a, b); }
        }
        __syncwarp();
    }
    if (FUSED) {
        for (int o = 16; o > 0; o >>= 1) visited += __shfl_down_sync(FULLM, visited, o);
        if (lane == 0 && visited) atomicAdd(out.counts + 3, visited);
    }
}

// totalcollisions_native (qtree_functions.jl:111-131): labels whose top kernel covers (L,1,1), in order
__global__ void k_native_filter(const LevelMeta* __restrict__ lm, int L, int nl, const int32_t* __restrict__ labels,
                                int32_t* __restrict__ flt, int* __restrict__ m_out) {
    __shared__ int wcnt[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j0 = 0; j0 < nl; j0 += blockDim.x) {
        int j = j0 + threadIdx.x;
        int o = 0; bool keep = false;
        if (j < nl) {
            o = labels ? labels[j] - 1 : j;
            LevelMeta m = lm[(size_t)o * L + (L - 1)];
            int rr = 1 - m.rs - 1, cc = 1 - m.cs - 1;
            keep = (unsigned)rr < (unsigned)lm_kh(m) && (unsigned)cc < (unsigned)lm_kw(m);
        }
        unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wcnt[w] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int i = 0; i < w; i++) off += wcnt[i];
        if (keep) flt[off + __popc(bal & ((1u << lane) - 1))] = o + 1;
        __syncthreads();
        if (threadIdx.x == 0) { int s = 0; for (int i = 0; i < nw; i++) s += wcnt[i]; base += s; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *m_out = base;
}
// pairs (labels[i], labels[j]) for i in 1:l for j in l:-1:i+1 (:120), start node (L,1,1)
__global__ void k_native_pairs(const int32_t* __restrict__ flt, const int* __restrict__ m_in, int nl, int L, Item16* __restrict__ items,
                               int64_t cap, unsigned long long* __restrict__ counts) {
    int m = *m_in;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) counts[0] = (unsigned long long)m * (unsigned long long)(m - 1 > 0 ? m - 1 : 0) / 2;
    int64_t i = t / nl, j = t - i * nl;
    if (i < j && j < m) {
        int64_t idx = i * (2 * (int64_t)m - i - 1) / 2 + (m - 1 - j);
        if (idx < cap) items[idx] = make_item(flt[i], flt[j], L, 1, 1);
    }
}

// =============================================================================================
// (3) narrow phase: _collision_randbfs (qtree_functions.jl:21-42) in canonical child order
// =============================================================================================
// The FIFO queue of the reference visits the tree level by level, so the BFS is run as a
// level-synchronous frontier.  The frontier (nodes whose code is MIX in both trees) is kept in BFS
// order; the 4*count child tests of one level are spread over the lanes of a group in that same
// order; a ballot finds the first FULL code (= the node the sequential BFS returns) and an ordered
// ballot/popc compaction appends the MIX children to the next frontier.
// Frontier entries are (alpha,beta) offsets below the start node: a = at.a*2^d - alpha.
#define NP_THREADS 128

struct NarrowArgs {
    const uint32_t* words; const LevelMeta* lm; int L;
    const Item16* items; const unsigned long long* n_items_dev; int64_t n_items_host; // explicit item list (k_narrow_items); device count if non-null
    int4* res;                       // per item (l,a,b,0); l < 0 = miss (NULL: not wanted)
    Item16* hits; int64_t hits_cap; unsigned long long* nhits; // hits (l >= 0) appended here (NULL: not wanted)
    unsigned long long* counts;      // [2] overflow count, [3] node tests
    SurvRec* overflow;               // pairs that need the big path
    int64_t overflow_cap;
    SurvRec* surv; int64_t surv_cap; unsigned long long* nsurv; // word-parallel part -> frontier kernel
    unsigned long long* cursor;      // frontier-kernel work cursor into surv (zeroed by the kernel that fills surv)
    int* status;                     // latency path: STF_ST_RETRY_GROW when the survivor list was too small
};

// Word-parallel part over a compact item list (the candidate list of the broad phase, or an explicit list of
// stf_collision_bfs / stf_collide_items / totalcollisions_native): nw_depth1..3, one THREAD per pair and depth.  The pairs
// that go one level deeper are COMPACTED per warp between the depths (two small shared-memory queues per warp, ballot
// ranks, no block barrier): a warp runs depth 1 on 32 fresh pairs, and whenever 32 pairs wait for depth 2 (or depth 3)
// it runs that depth on them with every lane busy — instead of dragging its few deep pairs through all three depths.
// Warps never wait for each other, so the gathers of one warp hide behind the arithmetic of the others.  Pairs still
// undecided after depth 3 go on to the frontier kernel below with their depth-3 frontier.
#define NW_THREADS 256
#define NW_QCAP 64
struct NwQEntry { Item16 item; uint32_t mix; int vis; uint32_t it; uint32_t pad; };
__global__ void __launch_bounds__(NW_THREADS) k_narrow_items(NarrowArgs A) {
    pdl_sync(); // programmatic dependent launch: scheduled while the previous kernel drains; nothing global before this
    __shared__ NwQEntry s_q[NW_THREADS / 32][2][NW_QCAP];
    const unsigned FULLM = 0xffffffffu;
    int64_t n = A.n_items_dev ? (int64_t)*A.n_items_dev : A.n_items_host;
    if (A.n_items_dev && n > A.n_items_host) n = A.n_items_host; // clamp to buffer capacity
    unsigned long long visited = 0;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0) *A.cursor = 0ull;
    NwQEntry* q2 = s_q[wib][0]; NwQEntry* q3 = s_q[wib][1];
    int n2 = 0, n3 = 0; // queue fills (warp-uniform)
    auto finish = [&](uint32_t it, const Item16& item, const NwResult& R) {
        visited += (unsigned long long)R.visited;
        if (R.undecided) {
            unsigned long long s_ = atomicAdd(A.nsurv, 1ull);
            if ((int64_t)s_ < A.surv_cap) { SurvRec sr; sr.item = item; sr.front = R.front; sr.depth = 3u; sr.it = it; A.surv[s_] = sr; }
            return;
        }
        if (A.res) A.res[it] = make_int4(R.l, R.a, R.b, 0);
        if (A.hits && R.l >= 0) {
            unsigned long long s_ = atomicAdd(A.nhits, 1ull);
            if ((int64_t)s_ < A.hits_cap) A.hits[s_] = make_item(item.i1, item_i2(item), R.l, R.a, R.b);
        }
    };
    const int64_t nwarps = (int64_t)gridDim.x * (NW_THREADS / 32);
    int64_t base = ((int64_t)blockIdx.x * (NW_THREADS / 32) + wib) * 32;
    for (;;) {
        const bool more = base < n;
        if (n3 >= 32 || (!more && n2 == 0 && n3 > 0)) { // ---- depth 3 on (up to) 32 queued pairs
            const int take = min(n3, 32), e = n3 - take + lane;
            if (lane < take) {
                const NwQEntry q = q3[e];
                const int l = item_l(q.item), a = q.item.a, b = q.item.b;
                const LevelMeta m1 = ldmeta<false>(A.lm + (size_t)(q.item.i1 - 1) * A.L + (l - 4)), m2 = ldmeta<false>(A.lm + (size_t)(item_i2(q.item) - 1) * A.L + (l - 4));
                uint32_t pa[4], pb[4];
                nw_block3<false>(A.words, m1, a, b, pa); nw_block3<false>(A.words, m2, a, b, pb);
                NwResult R; R.visited = q.vis; R.undecided = 0; R.front = 0;
                nw_depth3(pa, pb, q.mix, l, a, b, R);
                finish(q.it, q.item, R);
            }
            n3 -= take;
            __syncwarp();
            continue;
        }
        if (n2 >= 32 || (!more && n2 > 0)) { // ---- depth 2
            const int take = min(n2, 32), e = n2 - take + lane;
            bool deeper = false; NwQEntry q = {}; 
            if (lane < take) {
                q = q2[e];
                const int l = item_l(q.item), a = q.item.a, b = q.item.b;
                const LevelMeta m1 = ldmeta<false>(A.lm + (size_t)(q.item.i1 - 1) * A.L + (l - 3)), m2 = ldmeta<false>(A.lm + (size_t)(item_i2(q.item) - 1) * A.L + (l - 3));
                NwResult R; R.visited = q.vis; R.undecided = 0; R.front = 0; uint32_t mix;
                deeper = nw_depth2(nw_block2<false>(A.words, m1, a, b), nw_block2<false>(A.words, m2, a, b), q.mix, l, a, b, R, &mix) != 0;
                if (deeper) { q.mix = mix; q.vis = R.visited; } else finish(q.it, q.item, R);
            }
            n2 -= take;
            const unsigned dm = __ballot_sync(FULLM, deeper);
            __syncwarp(); // q2 reads of this round are done before anything is pushed again
            if (deeper) q3[n3 + __popc(dm & ((1u << lane) - 1u))] = q;
            n3 += __popc(dm);
            __syncwarp();
            continue;
        }
        if (!more) break;
        { // ---- depth 1 on 32 fresh pairs
            const int64_t it = base + lane;
            bool deeper = false; NwQEntry q = {};
            if (it < n) {
                q.item = A.items[it]; q.it = (uint32_t)it;
                const int l = item_l(q.item), a = q.item.a, b = q.item.b;
                const LevelMeta m1 = ldmeta<false>(A.lm + (size_t)(q.item.i1 - 1) * A.L + (l - 2)), m2 = ldmeta<false>(A.lm + (size_t)(item_i2(q.item) - 1) * A.L + (l - 2));
                NwResult R; uint32_t mix;
                deeper = nw_depth1(nw_block1<false>(A.words, m1, a, b), nw_block1<false>(A.words, m2, a, b), l, a, b, R, &mix) != 0;
                if (deeper) { q.mix = mix; q.vis = R.visited; } else finish(q.it, q.item, R);
            }
            const unsigned dm = __ballot_sync(FULLM, deeper);
            if (deeper) q2[n2 + __popc(dm & ((1u << lane) - 1u))] = q;
            n2 += __popc(dm);
            __syncwarp();
            base += nwarps * 32;
        }
    }
    for (int o = 16; o > 0; o >>= 1) visited += __shfl_down_sync(FULLM, visited, o);
    if (lane == 0 && visited) atomicAdd(A.counts + 3, visited);
}

// Frontier kernel, one G-lane GROUP per pair the word-parallel part left undecided, G / 4 frontier nodes (their four
// children each) per group and loop iteration.  G = 4 for batches (many pairs: throughput, eight pairs per warp); G = 32
// for single scenes (few pairs, each a long dependent chain: a whole warp expands eight frontier nodes per round trip, and
// the frontier buffer is eight times larger, so fewer pairs fall through to the CTA-per-pair path).
// The groups of a warp run ONE converged loop: a group that finishes its pair takes the next one from the warp's reserved
// chunk of the survivor list in the same iteration, so trip-count differences between pairs never idle lanes, and the
// votes are full-warp ballots.  Lanes are ordered (frontier node, child number) = BFS order: the first FULL code among
// the group's lanes is the node the sequential BFS returns, and the ballot/popc compaction keeps the next frontier in order.
template <int G> struct NpCfg { static constexpr int FCAP = G == 4 ? 64 : 512; };
template <int G>
__global__ void __launch_bounds__(NP_THREADS) k_collide_small(NarrowArgs A) {
    pdl_sync(); // programmatic dependent launch: scheduled while the previous kernel drains; nothing global before this
    constexpr int FCAP = NpCfg<G>::FCAP, NPI = G / 4; // frontier nodes per iteration
    __shared__ uint32_t fr[NP_THREADS / G][2][FCAP];
    const unsigned FULLM = 0xffffffffu;
    const int lane = threadIdx.x & 31, gl_ = lane % G, g0 = lane - gl_, gl = gl_ & 3, slot = gl_ >> 2;
    const int gib = threadIdx.x / G;
    const unsigned gmask = G == 32 ? FULLM : ((1u << G) - 1u);
    const unsigned below = G == 32 ? 0u : ((1u << g0) - 1u); // lanes of the groups before mine
    int64_t n = (int64_t)*A.nsurv;          // only the pairs the word-parallel part could not decide
    if (n > A.surv_cap) { // the list was too small: the round is re-run with a larger one (latency path) / the host re-runs
        if (A.status && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(A.status, 1 /* STF_ST_RETRY_GROW */);
        n = A.surv_cap;
    }
    const int64_t nwarps = (int64_t)gridDim.x * (NP_THREADS / 32);
    const int chunk = G == 32 ? 1 : (int)min((int64_t)64, max((int64_t)8, ((n / nwarps) + 7) & ~(int64_t)7));
    // the warp's chunk of the survivor list (warp-uniform): the first one is static, later ones come from the cursor
    int64_t cpos = ((int64_t)blockIdx.x * (NP_THREADS / 32) + (threadIdx.x >> 5)) * chunk, cend = min(cpos + chunk, n);
    bool exhausted = cpos >= n;
    if (exhausted) cpos = cend = 0;
    // group state (identical in the lanes of a group)
    bool active = false, newlevel = false;
    uint32_t it = 0; int lev0 = 0; // result index (explicit lists) and start level of the pair
    int i1 = 0, i2 = 0, lev = 0, ata = 0, atb = 0, cnt = 0, cur = 0, depth = 0, q = 0, ncnt = 0;
    LevelMeta m1 = {}, m2 = {};
    unsigned long long visited = 0;
    for (;;) {
        unsigned need = __ballot_sync(FULLM, !active);
        if (need) {
            if (cpos >= cend && !exhausted) {
                unsigned long long c0 = 0;
                if (lane == 0) c0 = atomicAdd(A.cursor, (unsigned long long)chunk);
                c0 = __shfl_sync(FULLM, c0, 0);
                cpos = (int64_t)c0 + nwarps * chunk; cend = min(cpos + chunk, n);
                if (cpos >= n) { exhausted = true; cpos = cend = 0; }
            }
            if (!active) {
                int64_t j = cpos + (__popc(need & below) / G);
                if (j < cend) {
                    const SurvRec sr = A.surv[j]; // the word-parallel part already tested `depth` levels: `front` = its last frontier
                    it = sr.it; lev0 = item_l(sr.item);
                    i1 = sr.item.i1; i2 = item_i2(sr.item); depth = (int)sr.depth; lev = lev0 - depth; ata = sr.item.a; atb = sr.item.b;
                    cnt = __popcll(sr.front); cur = 0; q = 0; ncnt = 0; newlevel = true; active = true;
                    uint64_t f = sr.front; int idx = 0; // Morton order = BFS order: set bit number idx is frontier entry idx
                    while (f) {
                        const int k = __ffsll((long long)f) - 1; f &= f - 1;
                        if (idx % G == gl_) { int al, be; nw_unkey((uint32_t)k, &al, &be); fr[gib][0][idx] = ((uint32_t)al << 16) | (uint32_t)be; }
                        idx++;
                    }
                }
            }
            cpos = min(cpos + (__popc(need) / G), cend);
            if (!__any_sync(FULLM, active)) { if (exhausted) break; continue; }
        }
        __syncwarp(); // frontier writes of the previous iteration / the start frontier
        if (newlevel) {
            m1 = A.lm[(size_t)(i1 - 1) * A.L + (lev - 2)]; m2 = A.lm[(size_t)(i2 - 1) * A.L + (lev - 2)];
            newlevel = false;
        }
        uint32_t code = 0, ab = 0; int a = 0, b = 0;
        const bool valid = active && q + slot < cnt;
        if (valid) {
            uint32_t node = fr[gib][cur][q + slot];
            uint32_t al = ((node >> 16) << 1) + (gl & 1), be = ((node & 0xffffu) << 1) + (gl >> 1);
            ab = (al << 16) | be;
            a = (ata << (depth + 1)) - (int)al; b = (atb << (depth + 1)) - (int)be;
            code = node_code<false>(A.words, m1, a, b) & node_code<false>(A.words, m2, a, b);
        }
        const unsigned hm = (__ballot_sync(FULLM, code == CODE_FULL) >> g0) & gmask;
        if (valid && (!hm || gl_ < __ffs(hm))) visited++; // the sequential BFS stops at the first FULL child
        const bool push = code == CODE_MIX && lev - 1 > 1;
        const unsigned pm = (__ballot_sync(FULLM, push) >> g0) & gmask;
        const int src = g0 + (hm ? __ffs(hm) - 1 : 0);
        const int ha = __shfl_sync(FULLM, a, src), hb = __shfl_sync(FULLM, b, src);
        if (active) {
            int4 result = make_int4(0, 0, 0, 0); bool done = false, overflow = false;
            if (hm) { result = make_int4(lev - 1, ha, hb, 0); done = true; }
            else {
                int np = __popc(pm);
                if (ncnt + np > FCAP) { overflow = true; done = true; }
                else {
                    if (push) fr[gib][cur ^ 1][ncnt + __popc(pm & ((1u << gl_) - 1u))] = ab;
                    ncnt += np; q += NPI;
                    if (q >= cnt) { // level finished
                        if (ncnt == 0) { // miss: -(level) and coordinates of the last node popped (:41)
                            uint32_t node = fr[gib][cur][cnt - 1];
                            result = make_int4(-lev, (ata << depth) - (int)(node >> 16), (atb << depth) - (int)(node & 0xffffu), 0);
                            done = true;
                        } else { cur ^= 1; cnt = ncnt; ncnt = 0; q = 0; lev -= 1; depth += 1; newlevel = true; }
                    }
                }
            }
            if (done) {
                if (gl_ == 0) {
                    if (overflow) {
                        unsigned long long s = atomicAdd(A.counts + 2, 1ull);
                        if ((int64_t)s < A.overflow_cap) { SurvRec sr; sr.item = make_item(i1, i2, lev0, ata, atb); sr.front = 0; sr.depth = 0u; sr.it = it; A.overflow[s] = sr; }
                        result = make_int4(-1, 0, 0, -1);
                    }
                    if (A.res && !overflow) A.res[it] = result;
                    if (A.hits && result.x >= 0) { // cp[1] >= 0  qtree_functions.jl:104-106
                        unsigned long long s = atomicAdd(A.nhits, 1ull);
                        if ((int64_t)s < A.hits_cap) A.hits[s] = make_item(i1, i2, result.x, result.y, result.z);
                    }
                }
                active = false;
            }
        }
    }
    // node-test counter (algorithmic bytes of the roofline)
    for (int o = 16; o > 0; o >>= 1) visited += __shfl_down_sync(0xffffffffu, visited, o);
    if ((threadIdx.x & 31) == 0 && visited) atomicAdd(A.counts + 3, visited);
}

// large-frontier path: one CTA per overflowed item, frontier double-buffered in global scratch
#define NPB_THREADS 256
__global__ void __launch_bounds__(NPB_THREADS) k_collide_big(NarrowArgs A, uint32_t* __restrict__ scratch, int64_t cap_per_cta, int* __restrict__ err) {
    pdl_sync(); // programmatic dependent launch: scheduled while the previous kernel drains; nothing global before this
    __shared__ int w_push[NPB_THREADS / 32];
    __shared__ int w_hit[NPB_THREADS / 32];
    __shared__ int4 s_result;
    int64_t nov = (int64_t)A.counts[2];
    if (nov > A.overflow_cap) nov = A.overflow_cap;
    uint32_t* f0 = scratch + (size_t)blockIdx.x * 2 * cap_per_cta;
    uint32_t* fbuf[2] = { f0, f0 + cap_per_cta };
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long visited = 0;
    for (int64_t ov = blockIdx.x; ov < nov; ov += gridDim.x) { // (restarts from the start node: the nodes the small path tested are counted twice)
        const SurvRec sr = A.overflow[ov];
        const uint32_t it = sr.it;
        const Item16 item = sr.item;
        int o1 = item.i1 - 1, o2 = item_i2(item) - 1;
        int lev = item_l(item), ata = item.a, atb = item.b;
        int64_t cnt = 1; int cur = 0, depth = 0;
        if (threadIdx.x == 0) fbuf[0][0] = 0;
        __syncthreads();
        int4 result = make_int4(-lev, ata, atb, 0);
        for (;;) {
            LevelMeta m1 = A.lm[(size_t)o1 * A.L + (lev - 2)], m2 = A.lm[(size_t)o2 * A.L + (lev - 2)];
            int basea = ata << (depth + 1), baseb = atb << (depth + 1);
            int64_t ncnt = 0; int done = 0;
            int64_t ntest = 4 * cnt;
            for (int64_t t0 = 0; t0 < ntest; t0 += NPB_THREADS) {
                int64_t t = t0 + threadIdx.x;
                bool valid = t < ntest;
                uint32_t code = 0, ab = 0; int a = 0, b = 0;
                if (valid) {
                    uint32_t node = fbuf[cur][t >> 2]; int cn = (int)(t & 3);
                    uint32_t al = ((node >> 16) << 1) + (cn & 1), be = ((node & 0xffffu) << 1) + (cn >> 1);
                    ab = (al << 16) | be;
                    a = basea - (int)al; b = baseb - (int)be;
                    code = node_code<false>(A.words, m1, a, b) & node_code<false>(A.words, m2, a, b);
                }
                visited += valid;
                unsigned hm = __ballot_sync(0xffffffffu, valid && code == CODE_FULL);
                bool push = valid && code == CODE_MIX && lev - 1 > 1;
                unsigned pm = __ballot_sync(0xffffffffu, push);
                if (lane == 0) { w_hit[w] = hm ? __ffs(hm) - 1 : -1; w_push[w] = __popc(pm); }
                __syncthreads();
                int first_w = -1; int off = 0, tot = 0;
                for (int i = 0; i < NPB_THREADS / 32; i++) {
                    if (first_w < 0 && w_hit[i] >= 0) first_w = i;
                    if (i < w) off += w_push[i];
                    tot += w_push[i];
                }
                if (first_w >= 0) {
                    if (w == first_w && lane == w_hit[w]) s_result = make_int4(lev - 1, a, b, 0);
                    done = 1;
                } else {
                    if (ncnt + tot > cap_per_cta) { if (threadIdx.x == 0) atomicOr(err, 4); done = 2; }
                    else if (push) fbuf[cur ^ 1][ncnt + off + __popc(pm & ((1u << lane) - 1))] = ab;
                    ncnt += tot;
                }
                __syncthreads();
                if (done) break;
            }
            if (done == 1) { result = s_result; break; }
            if (done == 2) { result = make_int4(-1, 0, 0, -2); break; }
            if (ncnt == 0) {
                uint32_t node = fbuf[cur][cnt - 1];
                result = make_int4(-lev, (ata << depth) - (int)(node >> 16), (atb << depth) - (int)(node & 0xffffu), 0);
                break;
            }
            __syncthreads();
            cur ^= 1; cnt = ncnt; lev -= 1; depth += 1;
        }
        if (threadIdx.x == 0) {
            if (A.res) A.res[it] = result;
            if (A.hits && result.x >= 0) {
                unsigned long long s = atomicAdd(A.nhits, 1ull);
                if ((int64_t)s < A.hits_cap) A.hits[s] = make_item(item.i1, item_i2(item), result.x, result.y, result.z);
            }
        }
        __syncthreads();
    }
    for (int o = 16; o > 0; o >>= 1) visited += __shfl_down_sync(0xffffffffu, visited, o);
    if (lane == 0 && visited) atomicAdd(A.counts + 3, visited);
}

// collision_dfs (qtree_functions.jl:2-20): one thread per pair, explicit stack
__global__ void k_collision_dfs(const uint32_t* __restrict__ words, const LevelMeta* __restrict__ lm, int L, const Item16* __restrict__ items, int64_t n, int4* __restrict__ res) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Item16 item = items[i];
    int o1 = item.i1 - 1, o2 = item_i2(item) - 1;
    int sa[STF_MAX_LEVELS + 1], sb[STF_MAX_LEVELS + 1], sc[STF_MAX_LEVELS + 1];
    int l = item_l(item), top = l;
    sa[l] = item.a; sb[l] = item.b; sc[l] = -1;
    int4 r = make_int4(-l, item.a, item.b, 0);
    while (l <= top) {
        if (sc[l] < 0) { // entering node (l, sa[l], sb[l])
            uint32_t code = node_code<false>(words, lm[(size_t)o1 * L + (l - 1)], sa[l], sb[l]) & node_code<false>(words, lm[(size_t)o2 * L + (l - 1)], sa[l], sb[l]);
            if (code == CODE_FULL) { r = make_int4(l, sa[l], sb[l], 0); break; }
            r = make_int4(-l, sa[l], sb[l], 0);
            if (code == CODE_MIX && l > 1) sc[l] = 0; else { l++; continue; }
        }
        if (sc[l] < 4) { int cn = sc[l]++; sa[l - 1] = 2 * sa[l] - (cn & 1); sb[l - 1] = 2 * sb[l] - (cn >> 1); sc[l - 1] = -1; l--; }
        else l++;
    }
    res[i] = r;
}

// =============================================================================================
// result ordering: sort!(r) and unique!(minmax∘first, ·)  (qtree_functions.jl:252)
// =============================================================================================
__global__ void k_hit_keys(const Item16* __restrict__ hits, int64_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    pdl_sync();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = ((uint64_t)(uint32_t)hits[i].i1 << 32) | (uint32_t)item_i2(hits[i]); vals[i] = (uint32_t)i; }
}
__global__ void k_gather_hits(const Item16* __restrict__ hits, const uint32_t* __restrict__ vals, int64_t n, Item16* __restrict__ out) {
    pdl_sync();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = hits[vals[i]];
}
__device__ __forceinline__ bool lab_less(const Item16& x, const Item16& y) {
    int lx = item_l(x), ly = item_l(y);
    if (lx != ly) return lx < ly; if (x.a != y.a) return x.a < y.a; return x.b < y.b;
}
// hits of one ordered pair (i1,i2) are adjacent after the radix sort; order them by (l,a,b) and mark what
// unique! keeps: the first entry of a run, unless the swapped pair (i2,i1) sorts earlier (i1 > i2 and it exists)
__global__ void k_fix_ties(Item16* __restrict__ hits, const uint64_t* __restrict__ keys, int64_t n, int unique, uint32_t* __restrict__ keep) {
    pdl_sync();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = keys[i];
    bool head = (i == 0) || keys[i - 1] != key;
    uint32_t kp = 1;
    if (unique) {
        kp = head;
        uint32_t i1 = (uint32_t)(key >> 32), i2 = (uint32_t)key;
        if (kp && i1 > i2) {
            uint64_t sw = ((uint64_t)i2 << 32) | i1;
            int64_t p = lower_bound_u64(keys, n, sw);
            if (p < n && keys[p] == sw) kp = 0;
        }
    }
    keep[i] = kp;
    if (!head) return;
    int64_t e = i + 1; while (e < n && keys[e] == key) e++;
    for (int64_t x = i + 1; x < e; x++) { // insertion sort of a (short) run
        Item16 v = hits[x]; int64_t y = x - 1;
        while (y >= i && lab_less(v, hits[y])) { hits[y + 1] = hits[y]; y--; }
        hits[y + 1] = v;
    }
}
__global__ void k_export_hits(const Item16* __restrict__ hits, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos, int64_t n, int32_t* __restrict__ out /*5 ints each*/) {
    pdl_sync();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    int32_t* o = out + 5 * (size_t)pos[i];
    Item16 h = hits[i];
    o[0] = h.i1; o[1] = item_i2(h); o[2] = item_l(h); o[3] = h.a; o[4] = h.b;
}
__global__ void k_import_items(const int32_t* __restrict__ in, int64_t n, Item16* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const int32_t* p = in + 5 * (size_t)i; out[i] = make_item(p[0], p[1], p[2], p[3], p[4]); }
}
__global__ void k_export_results(const Item16* __restrict__ items, const int4* __restrict__ res, int64_t n, int32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t* o = out + 5 * (size_t)i;
    o[0] = items[i].i1; o[1] = item_i2(items[i]); o[2] = res[i].x; o[3] = res[i].y; o[4] = res[i].z;
}
__global__ void k_export_items(const Item16* __restrict__ items, int64_t n, int32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t* o = out + 5 * (size_t)i;
    o[0] = items[i].i1; o[1] = item_i2(items[i]); o[2] = item_l(items[i]); o[3] = items[i].a; o[4] = items[i].b;
}
__global__ void k_export_cells(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const unsigned long long* __restrict__ n_dev, int64_t cap, int32_t* __restrict__ out) {
    int64_t n = min((int64_t)*n_dev, cap);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int l, a, b; cell_unkey(keys[i], &l, &a, &b);
    int32_t* o = out + 4 * (size_t)i;
    o[0] = l; o[1] = a; o[2] = b; o[3] = (int32_t)vals[i];
}

// ---- item-parallel sharding: the exchange buffer of one rank = [header: count][count hit records] padded to `stride`
__global__ void k_shard_pack(const Item16* __restrict__ hits, int64_t n, Item16* __restrict__ send, int64_t stride) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { Item16 h; h.i1 = (int32_t)(n & 0x7fffffff); h.i2l = (int32_t)(n >> 31); h.a = 0; h.b = 0; send[0] = h; }
    if (i < n && i + 1 < stride) send[i + 1] = hits[i];
}
// gathered segments of all ranks -> one compact hit list (rank order; the canonical order comes from the sort after)
__global__ void k_shard_unpack(const Item16* __restrict__ recv, int nranks, int64_t stride, Item16* __restrict__ hits, int64_t cap, unsigned long long* __restrict__ nhits) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int s = (int)(i / stride); int64_t k = i - (int64_t)s * stride;
    if (s >= nranks || k == 0) { if (i == 0) { unsigned long long t = 0; for (int q = 0; q < nranks; q++) { Item16 h = recv[(int64_t)q * stride]; t += (unsigned long long)(uint32_t)h.i1 | ((unsigned long long)(uint32_t)h.i2l << 31); } *nhits = t; } return; }
    int64_t base = 0, mine = 0;
    for (int q = 0; q <= s; q++) { Item16 h = recv[(int64_t)q * stride]; int64_t c = (int64_t)(uint32_t)h.i1 | ((int64_t)(uint32_t)h.i2l << 31); if (q < s) base += c; else mine = c; }
    if (k - 1 < mine && base + k - 1 < cap) hits[base + k - 1] = recv[i];
}

// ---- item-parallel group of GPUs driven by ONE process (stf_group_*): device-side exchange of the local hit lists over
// NVLink peer memory.  After its share of the narrow phase, GPU `rank` copies its hits into segment `rank` of EVERY peer's
// gather buffer with plain peer stores, fences at system scope, and the last block publishes (count | cancel bit, epoch)
// to every peer's flag slots with a system-scope release.  k_shard_gather then spins on the flags of all peers
// (system-scope acquire: a device-side barrier, no host involvement), ORs the cancel bits into the local status and
// compacts the segments, in rank order, into the local hit list.  Every GPU ends up with the same union.
#define STF_MAX_GROUP 16
struct PeerSet { Item16* gather[STF_MAX_GROUP]; unsigned long long* flags[STF_MAX_GROUP]; int n; };
__global__ void __launch_bounds__(256) k_shard_push(const Item16* __restrict__ hits, const unsigned long long* __restrict__ nh_dev, int64_t seg_cap,
                                                    PeerSet P, int rank, unsigned long long epoch, const int* __restrict__ status,
                                                    const unsigned long long* __restrict__ n_list_dev, int64_t list_cap, unsigned int* __restrict__ blocks_done) {
    pdl_sync();
    const int64_t nh = (int64_t)*nh_dev;
    const int64_t n = min(nh, seg_cap);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const Item16 h = hits[i];
        for (int p = 0; p < P.n; p++) P.gather[p][(int64_t)rank * seg_cap + i] = h;
    }
    __threadfence_system();
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(blocks_done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *blocks_done = 0u;
    if ((int)threadIdx.x < P.n) {
        // cancel bit: this GPU's part of the round cannot be used (a list overflowed, too many hits): every peer must drop the round too
        const bool bad = *status != 0 || nh > seg_cap || (n_list_dev && (int64_t)*n_list_dev > list_cap);
        unsigned long long* f = P.flags[threadIdx.x] + 2 * rank;
        f[0] = (unsigned long long)n | (bad ? (1ull << 63) : 0ull);
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(f + 1), "l"(epoch) : "memory");
    }
}
// a GPU that cannot run its part of the round (an allocation failed on the host side) still answers: empty list + cancel bit,
// so that no peer waits for it forever
__global__ void k_shard_poison(PeerSet P, int rank, unsigned long long epoch) {
    if ((int)threadIdx.x < P.n) {
        unsigned long long* f = P.flags[threadIdx.x] + 2 * rank;
        f[0] = 1ull << 63;
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(f + 1), "l"(epoch) : "memory");
    }
}
__global__ void __launch_bounds__(256) k_shard_gather(const Item16* __restrict__ gather, const unsigned long long* __restrict__ flags, int nranks, int64_t seg_cap,
                                                      unsigned long long epoch, Item16* __restrict__ hits, int64_t hits_cap, unsigned long long* __restrict__ nhits, int* __restrict__ status) {
    pdl_sync();
    __shared__ long long s_cnt[STF_MAX_GROUP], s_base[STF_MAX_GROUP + 1];
    if ((int)threadIdx.x < nranks) { // device-side barrier: wait until peer `threadIdx.x` has published this round
        const unsigned long long* f = flags + 2 * threadIdx.x;
        unsigned long long e;
        do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(e) : "l"(f + 1) : "memory"); } while (e < epoch);
        unsigned long long v = *((volatile const unsigned long long*)f);
        s_cnt[threadIdx.x] = (long long)(v & 0x7FFFFFFFFFFFFFFFull);
        if ((v >> 63) && blockIdx.x == 0) atomicOr(status, 1 /* STF_ST_RETRY_GROW: drop the round everywhere */);
    }
    __syncthreads();
    if (threadIdx.x == 0) { long long b = 0; for (int r = 0; r < nranks; r++) { s_base[r] = b; b += s_cnt[r]; } s_base[nranks] = b; if (blockIdx.x == 0) *nhits = (unsigned long long)b; }
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)nranks * seg_cap; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / seg_cap); const int64_t k = i - (int64_t)r * seg_cap;
        if (k < s_cnt[r] && s_base[r] + k < hits_cap) hits[s_base[r] + k] = gather[i];
    }
}

// ---- pair pools of the pairwise trainers (filttrain!, fit.jl:235-266): every pair is tested from the root (L,1,1) without
// a bounds pre-check; the pairs whose result level is >= nearlevel survive, in pool order.
// all pairs (i, j), i < j, of objects 1..n in the reference's order `for i in 1:n for j in i+1:n` (fit.jl:279): one thread per
// (i, j) of the square; item index = i*(2n-i-1)/2 + (j-i-1) (0-based)
__global__ void k_allpairs_items(int n, int L, Item16* __restrict__ items) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t i = t / n, j = t - i * n;
    if (i < j && j < n) items[i * (2 * (int64_t)n - i - 1) / 2 + (j - i - 1)] = make_item((int)i + 1, (int)j + 1, L, 1, 1);
}
__global__ void k_pairs_items(const int32_t* __restrict__ pairs, int64_t n, int L, Item16* __restrict__ items) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) items[t] = make_item(pairs[2 * t], pairs[2 * t + 1], L, 1, 1);
}
// order-preserving compaction of the results with level >= nearlevel: count per block, (k_scan_u32 over the block counts), write
#define FP_BLOCK 256
__global__ void __launch_bounds__(FP_BLOCK) k_near_count(const int4* __restrict__ res, int64_t n, int nearlevel, uint32_t* __restrict__ bcount) {
    int64_t i = (int64_t)blockIdx.x * FP_BLOCK + threadIdx.x;
    bool keep = i < n && res[i].x >= nearlevel;
    int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) bcount[blockIdx.x] = (uint32_t)c;
}
__global__ void __launch_bounds__(FP_BLOCK) k_near_write(const Item16* __restrict__ items, const int4* __restrict__ res, int64_t n, int nearlevel,
                                                         const uint32_t* __restrict__ boff, int32_t* __restrict__ out5) {
    __shared__ int wcnt[FP_BLOCK / 32];
    int64_t i = (int64_t)blockIdx.x * FP_BLOCK + threadIdx.x;
    bool keep = i < n && res[i].x >= nearlevel;
    unsigned bal = __ballot_sync(0xffffffffu, keep);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wcnt[w] = __popc(bal);
    __syncthreads();
    if (!keep) return;
    int off = 0;
    for (int k = 0; k < w; k++) off += wcnt[k];
    int32_t* o = out5 + 5 * ((size_t)boff[blockIdx.x] + off + __popc(bal & ((1u << lane) - 1u)));
    o[0] = items[i].i1; o[1] = item_i2(items[i]); o[2] = res[i].x; o[3] = res[i].y; o[4] = res[i].z;
}
