// stf_step.cuh — subsystem (4): grad2d + step!/step_mask!/move! + optimiser + rebuild, for a whole
// collision list in ONE launch.  Replaces batchsteps!/step_index!/step!/step_mask!/move!
// (fit.jl:11-75), grad2d/intlog2 (fit_helper.jl:51-78) and Momentum/SGD (fit_helper.jl:13-38).
//
// Canonical mode: the result equals the sequential loop over the list in list order.  A step only
// touches its own two trees (the mask is read, never moved), so steps on disjoint objects commute;
// each step therefore waits only for the previous step of each of its movable objects.  Warps claim
// steps through an atomic ticket in list order, which guarantees that every step a warp waits for
// is already owned by a running warp (no deadlock).  All state of the moved trees is read and
// written with L2-scope (.cg) accesses, because another SM may have produced it microseconds ago.
#pragma once
#include "stf_common.cuh"
#include "stf_pyramid.cuh"
#include "../../include/stuffing_b200_rng.h"

struct OptState { int kind; double eta, rho; };

// endpoints: one per (step k, side s) with a movable object; key = object, stable-sorted so that the
// entries of one object are in list order.  prev[2k+s] = index of the previous step of that object (+1), 0 = none.
__global__ void k_step_endpoints(const Item16* __restrict__ hits, int64_t n, const ObjMeta* __restrict__ om, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    pdl_sync();
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * n) return;
    int64_t k = e >> 1; int s = (int)(e & 1);
    int o = (s ? item_i2(hits[k]) : hits[k].i1) - 1;
    keys[e] = om[o].is_mask ? STF_KEY_INVALID : (uint64_t)(uint32_t)o;
    vals[e] = (uint32_t)e;
}
__global__ void k_step_prev(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const unsigned long long* __restrict__ m_dev, int64_t m_cap, uint32_t* __restrict__ prev) {
    pdl_sync();
    int64_t m = m_dev ? min((int64_t)*m_dev, m_cap) : m_cap;
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    uint64_t key = keys[p];
    uint32_t e = vals[p];
    uint32_t pv = 0;
    if (key != STF_KEY_INVALID && p > 0 && keys[p - 1] == key) pv = (vals[p - 1] >> 1) + 1;
    prev[e] = pv;
}

// Dependency depth of every step by Jacobi relaxation: after i sweeps lvl[k] = min(depth[k], i), so ANY number of sweeps
// yields lvl[pred] <= lvl[k], and (lvl, k) is a topological order (a predecessor always has the smaller k).  Claiming
// tickets in that order puts all steps without a pending predecessor first: independent chains (e.g. the scenes of a
// batch) overlap instead of being walked one after the other.
__global__ void k_step_level(const uint32_t* __restrict__ prev, const uint32_t* __restrict__ lin, uint32_t* __restrict__ lout, int64_t n) {
    pdl_sync();
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t p0 = prev[2 * k], p1 = prev[2 * k + 1], l = 0;
    if (p0) l = lin[p0 - 1] + 1;
    if (p1) l = max(l, lin[p1 - 1] + 1);
    lout[k] = l;
}
// The same dependency depth, EXACT, by ONE CTA in one launch (mid-size lists: 16 relaxation launches cost ~50 us whatever
// the work).  A predecessor always has the smaller index, so the list is processed in chunks of 1024 steps in order: the
// depths of earlier chunks are final, the chunk's own are relaxed in shared memory until nothing changes (as many sweeps as
// the longest chain inside the chunk).  Depths are clamped to 63 (one 6-bit radix pass): (level, k) stays a topological order.
__global__ void __launch_bounds__(1024) k_step_depth_cta(const uint32_t* __restrict__ prev, uint32_t* __restrict__ lvl, int64_t n) {
    pdl_sync();
    __shared__ uint32_t s_l[1024];
    const int t = threadIdx.x;
    for (int64_t c0 = 0; c0 < n; c0 += 1024) {
        const int64_t k = c0 + t;
        uint32_t p0 = 0, p1 = 0, base = 0;
        if (k < n) {
            p0 = prev[2 * k]; p1 = prev[2 * k + 1];
            if (p0 && (int64_t)(p0 - 1) < c0) { base = min(63u, lvl[p0 - 1] + 1); p0 = 0; }  // predecessor in an earlier chunk: final
            if (p1 && (int64_t)(p1 - 1) < c0) { base = max(base, min(63u, lvl[p1 - 1] + 1)); p1 = 0; }
        }
        s_l[t] = base;
        __syncthreads();
        for (;;) {
            uint32_t l = base;
            if (p0) l = max(l, min(63u, s_l[p0 - 1 - c0] + 1));
            if (p1) l = max(l, min(63u, s_l[p1 - 1 - c0] + 1));
            const int changed = __syncthreads_or(l != s_l[t]);
            s_l[t] = l;
            __syncthreads();
            if (!changed) break;
        }
        if (k < n) lvl[k] = s_l[t];
        __syncthreads(); // (the next chunk reads lvl[] of this one from global memory)
    }
}
__global__ void k_step_level_keys(const uint32_t* __restrict__ lvl, int64_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    pdl_sync();
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { keys[k] = lvl[k]; vals[k] = (uint32_t)k; }
}

__device__ __forceinline__ int intlog2_dev(double x) { return (int)((__double_as_longlong(x) >> 52) - 1023); } // fit_helper.jl:74-78

struct StepArgs {
    uint32_t* words; LevelMeta* lm; const ObjMeta* om; int L;
    const Item16* hits; int64_t n; const unsigned long long* n_dev; // n_dev != NULL: the step count lives on the device
    const uint32_t* prev; uint32_t* done; unsigned long long* ticket;
    const uint32_t* order; // optional: ticket t runs step order[t] (a topological order that interleaves independent chains)
    double* vel; uint8_t* has_vel; uint8_t* moved_flag; int* dirty; int* built_from; // built_from[o] = (level the last rebuild of o started from) - 1
    OptState opt; uint64_t seed, iter;
    unsigned long long* counts; // [4] number of moves
    unsigned long long* dbg_t;  // optional timeline (4 x u64 per step) when dbg & 16
    int dbg;                    // STF_DEBUG_STEP experiments: 1 no wait, 2 no rebuild, 4 no fence (breaks parity)
};

// optimiser(t, Δ)  fit_helper.jl:19-35, fit.jl:25 — every lane computes it, lane `st` stores; (has, v0, v1) were prefetched
__device__ __forceinline__ void opt_apply_dev(const StepArgs& A, int o, double& d0, double& d1, int has, double v0, double v1, bool st) {
    if (A.opt.kind == 0) { d0 = d0 / 6.0; d1 = d1 / 6.0; return; }
    if (A.opt.kind == 1) { d0 = __dmul_rn(A.opt.eta, d0); d1 = __dmul_rn(A.opt.eta, d1); return; }
    if (!has) { v0 = d0; v1 = d1; if (st) __stcg(A.has_vel + o, (uint8_t)1); } // v = get!(velocity, x, Δ)
    double omr = __dsub_rn(1.0, A.opt.rho);
    v0 = __dadd_rn(__dmul_rn(A.opt.rho, v0), __dmul_rn(omr, d0)); // no FMA contraction: Julia does not fuse
    v1 = __dadd_rn(__dmul_rn(A.opt.rho, v1), __dmul_rn(omr, d1));
    if (st) { __stcg(A.vel + 2 * (size_t)o, v0); __stcg(A.vel + 2 * (size_t)o + 1, v1); }
    d0 = __dmul_rn(A.opt.eta, v0); d1 = __dmul_rn(A.opt.eta, v1);
}
// move!  fit.jl:11-23 — decides (level, s1, s2).  r_jit, r_d1, r_d2: the three draws of this
// tree (jitter test, jitter diagonal, avoid-rest diagonal), computed in parallel by other lanes
__device__ __forceinline__ double u64_to_unit(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }
__device__ __forceinline__ int3 move_decide(const StepArgs& A, double d0, double d1, uint64_t r_jit, uint64_t r_d1, uint64_t r_d2) {
    if (u64_to_unit(r_jit) < 0.1) {
        int j = (int)(r_d1 >> 62); // ((1,-1),(-1,1),(-1,-1),(1,1))
        d0 = __dadd_rn(d0, (j == 0 || j == 3) ? 1.0 : -1.0); d1 = __dadd_rn(d1, (j == 1 || j == 3) ? 1.0 : -1.0);
    }
    if (-1 < d0 && d0 < 1 && -1 < d1 && d1 < 1) {
        int j = (int)(r_d2 >> 62);
        d0 = (j == 0 || j == 3) ? 1.0 : -1.0; d1 = (j == 1 || j == 3) ? 1.0 : -1.0;
    }
    double dm = fmax(fabs(d0), fabs(d1));
    int u = intlog2_dev(dm);
    // trunc(Int, Δ) ÷ 2^u (fit.jl:22): |Δ| < 2^(u+1), so each component is -1, 0 or 1 — the leading binary digit
    double p = (double)(1ll << u);
    int s1 = d0 >= p ? 1 : (d0 <= -p ? -1 : 0), s2 = d1 >= p ? 1 : (d1 <= -p ? -1 : 0);
    return make_int3(min(A.L, 1 + u), s1, s2);
}
// ---- lazy batchsteps -------------------------------------------------------------------------------------------
// A move (shift!, qtrees.jl:267-275) translates levels <= lv and rebuilds levels > lv.  A pyramid that is consistent
// (every level = buildqtree! of the level below) stays consistent under shift!, so it is a pure function of the level-1
// raster and the per-level shifts.  Inside one batch a step therefore only updates the SHIFTS of the moved tree (all
// levels, closed form) and records dirty[o] = lowest level whose raster is still valid in memory; the few gradient
// nodes a later step of the same batch needs above that level are derived on the fly (grad_lazy), and every moved tree
// is materialised ONCE after the batch (k_materialize).  No rebuild sits on the step-to-step dependency chain.
// Per step, the dependent chain is: flag -> ONE round trip for everything that describes the two trees (each lane
// loads the LevelMeta of "its" level — lanes 0-15 tree 1, 16-31 tree 2 — plus the meta of the hit level, the dirty
// levels and both velocities, all as warp-broadcast loads) -> ONE round trip for the 2x8 gradient nodes -> two
// ballots (the gradient sums are popcounts) -> the scalar arithmetic of step!/move!, done redundantly by every lane so
// that nothing has to be broadcast -> stores, release fence, flag.  Everything that does not depend on tree state
// (the hit, the draws, the mass-ratio tests) is computed before the wait.
// MINB = resident CTAs per SM the register budget is sized for: 4 (120 registers, shortest dependent chain) for single
// scenes, 6 (80 registers) for batches of scenes, where independent chains fill the extra warps
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_batchsteps(StepArgs A) {
    pdl_sync(); // programmatic dependent launch: scheduled while the previous kernel drains; nothing global before this
    __shared__ uint32_t sbuf[4][2 * RB_HALF]; // only the rare eager materialisation uses it
    __shared__ uint64_t s_rnd[4][8];
    __shared__ int2 s_shift[4][32];
    const unsigned FULLM = 0xffffffffu;
    int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, half = lane >> 4, hl = lane & 15;
    const uint64_t hbase = stf_mix64(stf_mix64(A.seed ^ 0x5354554646494E47ull) + A.iter); // stf_rand_u64 without the per-draw part
    const int64_t nsteps = A.n_dev ? (int64_t)*A.n_dev : A.n;
    int k_round = 0;
    for (;;) {
        unsigned long long k = 0;
        if (A.dbg & 32) { if (k_round++) return; k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; } // timing experiment: no ticket
        else {
            if (lane == 0) k = atomicAdd(A.ticket, 1ull);
            k = __shfl_sync(FULLM, k, 0);
        }
        if ((int64_t)k >= nsteps) return;
        if (A.order) k = A.order[k];
        unsigned long long t_ready = 0, t_decided = 0;
        Item16 h = A.hits[k];
        int o1 = h.i1 - 1, o2 = item_i2(h) - 1;
        int l = item_l(h), a = h.a, b = h.b;
        ObjMeta om1 = A.om[o1], om2 = A.om[o2];
        uint32_t pv = lane < 2 ? A.prev[2 * k + lane] : 0u;
        // the 8 draws of this step (stuffing_b200_rng.h slots 0..7): lane j makes draw j, shared memory hands them round
        if (lane < 8) s_rnd[wib][lane] = stf_mix64(hbase + (k << 3) + (uint64_t)lane);
        __syncwarp();
        uint64_t R[8];
#pragma unroll
        for (int j = 0; j < 8; j++) R[j] = s_rnd[wib][j];
        bool m1 = om1.is_mask, m2 = om2.is_mask;
        bool move1 = false, move2 = false;
        if (!(m1 || m2)) { // step!  fit.jl:33-35
            double s1 = (double)(om1.kh1 + om1.kw1), s2 = (double)(om2.kh1 + om2.kw1);
            double ks1 = s1 * s1 * s1, ks2 = s2 * s2 * s2;
            move1 = u64_to_unit(R[0]) < ks2 / ks1;
            move2 = u64_to_unit(R[1]) < ks1 / ks2;
        }
        double ll = (double)(1ll << (l - 1));
        int oh = half ? o2 : o1, oo = half ? o1 : o2;
        if (pv && !(A.dbg & 1)) { // wait for the previous step of each movable object: relaxed polls (no L1 invalidation), then
            const uint32_t* f = A.done + (pv - 1); uint32_t v; // one acquire; every later read of tree state is an L2 (.cg) read
            do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory"); } while (v == 0);
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        }
        __syncwarp();
        long long s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0;
        if (A.dbg & 16) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ready)); asm volatile("mov.u64 %0, %%clock64;" : "=l"(s0)); }
        // ---- round trip 1: tree state
        LevelMeta mine = {};
        if (hl < A.L) mine = ldmeta<true>(A.lm + (size_t)oh * A.L + hl);
        int dmh = __ldcg(A.dirty + oh), dmo = __ldcg(A.dirty + oo); // levels > dirty are stale in memory
        int hasA = 0, hasB = 0; double2 vA = make_double2(0, 0), vB = make_double2(0, 0);
        if (A.opt.kind == 2) {
            hasA = __ldcg(A.has_vel + o1); hasB = __ldcg(A.has_vel + o2);
            vA = __ldcg(reinterpret_cast<const double2*>(A.vel + 2 * (size_t)o1));
            vB = __ldcg(reinterpret_cast<const double2*>(A.vel + 2 * (size_t)o2));
        }
        int dmA = half ? dmo : dmh, dmB = half ? dmh : dmo;
        s_shift[wib][lane] = make_int2(mine.rs, mine.cs);
        if (A.dbg & 16) asm volatile("mov.u64 %0, %%clock64;" : "=l"(s1) : "r"(mine.off), "r"(dmh), "r"(dmo), "d"(vB.y), "d"(vA.x), "r"(hasB));
        // ---- round trip 2: grad2d (fit_helper.jl:51-66).  A tree that moved earlier in this batch has stale rasters above
        // dirty[o]: its 3x3 window is derived from the lowest valid level (lazy window); the level metas it needs are the ones
        // the lanes already hold (shuffles, no further memory round trips), and the loads of both trees and of the direct
        // nodes are issued before anything is consumed.  The rarely executed parts are kept as ONE loop body each: a step on
        // the dependent chain usually meets this code cold in the instruction cache, so its size is latency.
        if (l - dmA > 3 || l - dmB > 3) { // window too wide: materialise that tree now (rare), then read it directly
#pragma unroll 1
            for (int t = 0; t < 2; t++) {
                int dmt = t ? dmB : dmA, ot = t ? o2 : o1;
                if (l - dmt <= 3) continue;
                LevelMeta tm = {};
                if (lane < A.L) tm = ldmeta<true>(A.lm + (size_t)ot * A.L + lane);
                warp_rebuild<true, true>(A.words, tm, A.L, dmt, lane, sbuf[wib]);
                if (lane == 0) A.built_from[ot] = dmt - 1;
                __syncwarp();
                if (t) dmB = A.L; else dmA = A.L;
            }
        }
        // every tree goes through the window code (k = 0 for a valid level: the 3 columns of the 3x3 window are read directly),
        // so the dependent chain never meets a code path that no earlier warp of the SM has executed
        uint64_t xA = 0, xB = 0;
#pragma unroll 1
        for (int t = 0; t < 2; t++) {
            const int mt = min(t ? dmB : dmA, l);
            LevelMeta mm = shfl_meta(mine, 16 * t + mt - 1);
            uint64_t x = lazy_window_load<true>(A.words, mm, l, a, b, mt, lane);
            if (t) xB = x; else xA = x;
        }
        int g1x = 0, g1y = 0, g2x = 0, g2y = 0;
        if (A.dbg & 16) asm volatile("mov.u64 %0, %%clock64;" : "=l"(s2) : "l"(xA), "l"(xB));
#pragma unroll 1
        for (int t = 0; t < 2; t++) {
            const int mt = min(t ? dmB : dmA, l);
            int2 g = lazy_window_reduce(t ? xB : xA, [&](int j) { return shfl_meta(mine, 16 * t + j - 1); }, l, a, b, mt, lane);
            if (t) { g2x = g.x; g2y = g.y; } else { g1x = g.x; g1y = g.y; }
        }
        // ---- step!/step_mask! (fit.jl:25-56): every lane runs the same scalar arithmetic
        int3 mvA = make_int3(0, 0, 0), mvB = make_int3(0, 0, 0); int oA = -1, oB = -1;
        bool st = lane == 0; // the lane that stores the optimiser state
        {
            double d1x = ll * g1x, d1y = ll * g1y, d2x = ll * g2x, d2y = ll * g2y;
            if (m1 || m2) { // step_mask!  fit.jl:49-56 (i1 == 1 first, then i2 == 1: fit.jl:60-63)
                int ot = m1 ? o2 : o1;
                double dx = m1 ? (d2x - d1x) / 2 : (d1x - d2x) / 2, dy = m1 ? (d2y - d1y) / 2 : (d1y - d2y) / 2;
                if (m1) opt_apply_dev(A, ot, dx, dy, hasB, vB.x, vB.y, st); else opt_apply_dev(A, ot, dx, dy, hasA, vA.x, vA.y, st);
                mvA = move_decide(A, dx, dy, R[2], R[3], R[4]); oA = ot;
            } else { // step!  fit.jl:25-48
                int slot = 2;
                if (move1) {
                    double dx = d1x, dy = d1y;
                    if (!move2) { dx -= d2x; dy -= d2y; }
                    opt_apply_dev(A, o1, dx, dy, hasA, vA.x, vA.y, st);
                    mvA = move_decide(A, dx, dy, R[2], R[3], R[4]); oA = o1; slot += 3;
                }
                if (move2) {
                    double dx = d2x, dy = d2y;
                    if (!move1) { dx -= d1x; dy -= d1y; }
                    opt_apply_dev(A, o2, dx, dy, hasB, vB.x, vB.y, st);
                    int3 mv = slot == 2 ? move_decide(A, dx, dy, R[2], R[3], R[4]) : move_decide(A, dx, dy, R[5], R[6], R[7]);
                    if (oA < 0) { mvA = mv; oA = o2; } else { mvB = mv; oB = o2; }
                }
            }
        }
        if (A.dbg & 16) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_decided)); asm volatile("mov.u64 %0, %%clock64;" : "=l"(s3) : "r"(mvA.x), "r"(mvB.x), "r"(oA), "r"(oB)); }
        // ---- shift! on the metas only (qtrees.jl:267-275); both halves at once, half h holds tree (h ? o2 : o1)
        bool isA = oA == oh, isB = oB == oh; // o1 != o2; a tree moves at most once per step
        bool moves = isA || isB;
        int3 mv = isA ? mvA : mvB;
        int lv = moves ? mv.x : 1;
        __syncwarp();
        int2 base = s_shift[wib][half * 16 + lv - 1];
        if (moves && hl < A.L) {
            if (hl < lv) { int f = 1 << (lv - 1 - hl); mine.rs += mv.y * f; mine.cs += mv.z * f; }
            else { // levels above lv: (new shift of level lv) ÷ 2^sh toward zero (qtrees.jl:225-226 composed)
                int sh = hl + 1 - lv, brs = base.x + mv.y, bcs = base.y + mv.z;
                mine.rs = brs >= 0 ? (brs >> sh) : -((-brs) >> sh); mine.cs = bcs >= 0 ? (bcs >> sh) : -((-bcs) >> sh);
            }
            stmeta<true>(A.lm + (size_t)oh * A.L + hl, mine);
        }
        int dcur = half ? dmB : dmA;
        if (hl == 0) { int nd = moves ? min(dcur, lv) : dcur; if (nd != dmh) __stcg(A.dirty + oh, nd); }
        if (A.dbg & 16) asm volatile("mov.u64 %0, %%clock64;" : "=l"(s4));
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        __syncwarp();
        if (A.dbg & 16) asm volatile("mov.u64 %0, %%clock64;" : "=l"(s5));
        if (lane == 0) {
            *((volatile uint32_t*)(A.done + k)) = 1u;
            if (A.dbg & 16) { unsigned long long t_end; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end)); A.dbg_t[4 * k] = (unsigned long long)(((s1 - s0) >> 4) & 0xfff) | ((unsigned long long)(((s2 - s1) >> 4) & 0xfff) << 12) | ((unsigned long long)(((s3 - s2) >> 4) & 0xfff) << 24) | ((unsigned long long)(((s4 - s3) >> 4) & 0xfff) << 36) | ((unsigned long long)(((s5 - s4) >> 4) & 0xfff) << 48); A.dbg_t[4 * k + 1] = t_ready; A.dbg_t[4 * k + 2] = t_decided; A.dbg_t[4 * k + 3] = t_end; }
        }
        if (hl == 0 && moves) { A.moved_flag[oh] = 1; atomicAdd(A.counts + 4, 1ull); } // off the chain: only read after the kernel
    }
}

// ---- scene-local batchsteps (batches of scenes) -------------------------------------------------------------------
// The same canonical result, organised around shared memory instead of global tickets and flags.  The scenes of a batch
// never interact, and when scene ids ascend with the labels the hits of a scene are CONTIGUOUS in the canonical order
// (k_scene_ranges).  A CTA claims a scene (one global ticket per scene, not per step), loads its hits, links and object
// sizes in one coalesced round trip, computes the exact dependency depth of every step (relaxation in shared memory) and
// runs the scene LEVEL BY LEVEL: the steps of one depth touch disjoint trees, so they run concurrently, one warp per step,
// and a CTA barrier separates the levels — no flag, no gpu-scope fence and no poll between dependent steps (that hop
// costs ~2 us per dependent step in k_batchsteps), no global depth sweeps and no sort of the steps by depth.  A scene
// with more than CAP hits is run in consecutive pieces of CAP steps by the same CTA (earlier pieces are complete).
__global__ void k_scene_ranges(const Item16* __restrict__ hits, int64_t n, const ObjMeta* __restrict__ om, uint32_t* __restrict__ rng) {
    pdl_sync();
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int sc = om[hits[k].i1 - 1].scene;
    if (k == 0 || om[hits[k - 1].i1 - 1].scene != sc) rng[2 * sc] = (uint32_t)k;
    if (k == n - 1 || om[hits[k + 1].i1 - 1].scene != sc) rng[2 * sc + 1] = (uint32_t)(k + 1);
}
__device__ __forceinline__ void step_core(const StepArgs& A, const unsigned long long k, const Item16 h, const int sz1, const int sz2,
                                          const uint64_t hbase, const int lane, uint64_t* __restrict__ s_rnd, int2* __restrict__ s_shift, uint32_t* __restrict__ s_sbuf, int* __restrict__ s_lock) {
    const int half = lane >> 4, hl = lane & 15;
    const int o1 = h.i1 - 1, o2 = item_i2(h) - 1;
    const int l = item_l(h), a = h.a, b = h.b;
    if (lane < 8) s_rnd[lane] = stf_mix64(hbase + (k << 3) + (uint64_t)lane);
    __syncwarp();
    uint64_t R[8];
#pragma unroll
    for (int j = 0; j < 8; j++) R[j] = s_rnd[j];
    const bool m1 = sz1 < 0, m2 = sz2 < 0;
    bool move1 = false, move2 = false;
    if (!(m1 || m2)) { // step!  fit.jl:33-35
        double s1 = (double)sz1, s2 = (double)sz2;
        double ks1 = s1 * s1 * s1, ks2 = s2 * s2 * s2;
        move1 = u64_to_unit(R[0]) < ks2 / ks1;
        move2 = u64_to_unit(R[1]) < ks1 / ks2;
    }
    const double ll = (double)(1ll << (l - 1));
    const int oh = half ? o2 : o1, oo = half ? o1 : o2;
    // ---- round trip 1: tree state
    LevelMeta mine = {};
    if (hl < A.L) mine = ldmeta<true>(A.lm + (size_t)oh * A.L + hl);
    int dmh = __ldcg(A.dirty + oh), dmo = __ldcg(A.dirty + oo); // levels > dirty are stale in memory
    int hasA = 0, hasB = 0; double2 vA = make_double2(0, 0), vB = make_double2(0, 0);
    if (A.opt.kind == 2) {
        hasA = __ldcg(A.has_vel + o1); hasB = __ldcg(A.has_vel + o2);
        vA = __ldcg(reinterpret_cast<const double2*>(A.vel + 2 * (size_t)o1));
        vB = __ldcg(reinterpret_cast<const double2*>(A.vel + 2 * (size_t)o2));
    }
    int dmA = half ? dmo : dmh, dmB = half ? dmh : dmo;
    s_shift[lane] = make_int2(mine.rs, mine.cs);
    // ---- round trip 2: grad2d (fit_helper.jl:51-66); stale trees through the lazy window (see k_batchsteps)
    if (l - dmA > 3 || l - dmB > 3) { // window too wide: materialise that tree now (rare); one scratch buffer per CTA
#pragma unroll 1
        for (int t = 0; t < 2; t++) {
            int dmt = t ? dmB : dmA, ot = t ? o2 : o1;
            if (l - dmt <= 3) continue;
            LevelMeta tm = {};
            if (lane < A.L) tm = ldmeta<true>(A.lm + (size_t)ot * A.L + lane);
            if (lane == 0) while (atomicCAS(s_lock, 0, 1) != 0) { }
            __syncwarp();
            warp_rebuild<true, true>(A.words, tm, A.L, dmt, lane, s_sbuf);
            __syncwarp();
            if (lane == 0) { A.built_from[ot] = dmt - 1; __threadfence_block(); atomicExch(s_lock, 0); }
            __syncwarp();
            if (t) dmB = A.L; else dmA = A.L;
        }
    }
    uint64_t xA = 0, xB = 0; // (every tree through the window code, k = 0 for a valid level: see k_batchsteps)
#pragma unroll 1
    for (int t = 0; t < 2; t++) {
        const int mt = min(t ? dmB : dmA, l);
        LevelMeta mm = shfl_meta(mine, 16 * t + mt - 1);
        uint64_t x = lazy_window_load<true>(A.words, mm, l, a, b, mt, lane);
        if (t) xB = x; else xA = x;
    }
    int g1x = 0, g1y = 0, g2x = 0, g2y = 0;
#pragma unroll 1
    for (int t = 0; t < 2; t++) {
        const int mt = min(t ? dmB : dmA, l);
        int2 g = lazy_window_reduce(t ? xB : xA, [&](int j) { return shfl_meta(mine, 16 * t + j - 1); }, l, a, b, mt, lane);
        if (t) { g2x = g.x; g2y = g.y; } else { g1x = g.x; g1y = g.y; }
    }
    // ---- step!/step_mask! (fit.jl:25-56): every lane runs the same scalar arithmetic
    int3 mvA = make_int3(0, 0, 0), mvB = make_int3(0, 0, 0); int oA = -1, oB = -1;
    const bool st = lane == 0; // the lane that stores the optimiser state
    {
        double d1x = ll * g1x, d1y = ll * g1y, d2x = ll * g2x, d2y = ll * g2y;
        if (m1 || m2) { // step_mask!  fit.jl:49-56 (i1 == 1 first, then i2 == 1: fit.jl:60-63)
            int ot = m1 ? o2 : o1;
            double dx = m1 ? (d2x - d1x) / 2 : (d1x - d2x) / 2, dy = m1 ? (d2y - d1y) / 2 : (d1y - d2y) / 2;
            if (m1) opt_apply_dev(A, ot, dx, dy, hasB, vB.x, vB.y, st); else opt_apply_dev(A, ot, dx, dy, hasA, vA.x, vA.y, st);
            mvA = move_decide(A, dx, dy, R[2], R[3], R[4]); oA = ot;
        } else { // step!  fit.jl:25-48
            int slot = 2;
            if (move1) {
                double dx = d1x, dy = d1y;
                if (!move2) { dx -= d2x; dy -= d2y; }
                opt_apply_dev(A, o1, dx, dy, hasA, vA.x, vA.y, st);
                mvA = move_decide(A, dx, dy, R[2], R[3], R[4]); oA = o1; slot += 3;
            }
            if (move2) {
                double dx = d2x, dy = d2y;
                if (!move1) { dx -= d1x; dy -= d1y; }
                opt_apply_dev(A, o2, dx, dy, hasB, vB.x, vB.y, st);
                int3 mv = slot == 2 ? move_decide(A, dx, dy, R[2], R[3], R[4]) : move_decide(A, dx, dy, R[5], R[6], R[7]);
                if (oA < 0) { mvA = mv; oA = o2; } else { mvB = mv; oB = o2; }
            }
        }
    }
    // ---- shift! on the metas only (qtrees.jl:267-275); both halves at once, half h holds tree (h ? o2 : o1)
    const bool isA = oA == oh, isB = oB == oh; // o1 != o2; a tree moves at most once per step
    const bool moves = isA || isB;
    const int3 mv = isA ? mvA : mvB;
    const int lv = moves ? mv.x : 1;
    __syncwarp();
    const int2 base = s_shift[half * 16 + lv - 1];
    if (moves && hl < A.L) {
        if (hl < lv) { int f = 1 << (lv - 1 - hl); mine.rs += mv.y * f; mine.cs += mv.z * f; }
        else { // levels above lv: (new shift of level lv) ÷ 2^sh toward zero (qtrees.jl:225-226 composed)
            int sh = hl + 1 - lv, brs = base.x + mv.y, bcs = base.y + mv.z;
            mine.rs = brs >= 0 ? (brs >> sh) : -((-brs) >> sh); mine.cs = bcs >= 0 ? (bcs >> sh) : -((-bcs) >> sh);
        }
        stmeta<true>(A.lm + (size_t)oh * A.L + hl, mine);
    }
    const int dcur = half ? dmB : dmA;
    if (hl == 0) { int nd = moves ? min(dcur, lv) : dcur; if (nd != dmh) __stcg(A.dirty + oh, nd); }
    if (hl == 0 && moves) { A.moved_flag[oh] = 1; atomicAdd(A.counts + 4, 1ull); } // only read after the kernel
    __syncwarp();
}

template <int NT, int CAP, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_batchsteps_scene(StepArgs A, const uint32_t* __restrict__ rng, int nscenes) {
    pdl_sync(); // programmatic dependent launch: nothing global before this
    constexpr int NW = NT / 32;
    const unsigned FULLM = 0xffffffffu;
    __shared__ Item16 s_hit[CAP];
    __shared__ int s_sz[2 * CAP];      // kh1 + kw1 of the object on each side of the hit; -1 = the mask
    __shared__ uint16_t s_p0[CAP], s_p1[CAP]; // links inside the piece: local index + 1, 0 = none (or already complete)
    __shared__ uint16_t s_depth[CAP], s_order[CAP];
    __shared__ int s_lstart[CAP + 1], s_cur[CAP];
    __shared__ uint64_t s_rnd[NW][8];
    __shared__ int2 s_shift[NW][32];
    __shared__ uint32_t s_sbuf[2 * RB_HALF];
    __shared__ int s_lock, s_nlev, s_scene;
    const int t = threadIdx.x, lane = t & 31, wib = t >> 5;
    const uint64_t hbase = stf_mix64(stf_mix64(A.seed ^ 0x5354554646494E47ull) + A.iter); // stf_rand_u64 without the per-draw part
    if (t == 0) s_lock = 0;
    for (;;) {
        __syncthreads(); // the previous scene's shared arrays are no longer read
        if (t == 0) s_scene = (int)atomicAdd(A.ticket, 1ull);
        __syncthreads();
        const int sc = s_scene;
        if (sc >= nscenes) return;
        const int64_t lo = rng[2 * sc], hi = rng[2 * sc + 1];
        for (int64_t c0 = lo; c0 < hi; c0 += CAP) {
            const int m = (int)min((int64_t)CAP, hi - c0);
            // ---- one coalesced round trip: hits, links, object sizes
            for (int i = t; i < m; i += NT) {
                const Item16 h = A.hits[c0 + i];
                const uint2 pv = *reinterpret_cast<const uint2*>(A.prev + 2 * (c0 + i));
                const ObjMeta a = A.om[h.i1 - 1], b = A.om[item_i2(h) - 1];
                s_hit[i] = h;
                s_sz[2 * i] = a.is_mask ? -1 : a.kh1 + a.kw1; s_sz[2 * i + 1] = b.is_mask ? -1 : b.kh1 + b.kw1;
                s_p0[i] = (pv.x && (int64_t)(pv.x - 1) >= c0) ? (uint16_t)(pv.x - c0) : (uint16_t)0;
                s_p1[i] = (pv.y && (int64_t)(pv.y - 1) >= c0) ? (uint16_t)(pv.y - c0) : (uint16_t)0;
                s_depth[i] = 0; s_cur[i] = 0;
            }
            for (int i = t; i <= CAP; i += NT) s_lstart[i] = 0;
            __syncthreads();
            // ---- exact depth inside the piece: relax until nothing changes (as many sweeps as the longest chain)
            for (;;) {
                int changed = 0;
                uint32_t nl[(CAP + NT - 1) / NT];
#pragma unroll
                for (int j = 0; j < (CAP + NT - 1) / NT; j++) {
                    const int i = t + j * NT;
                    uint32_t l = 0;
                    if (i < m) {
                        const int p0 = s_p0[i], p1 = s_p1[i];
                        if (p0) l = s_depth[p0 - 1] + 1u;
                        if (p1) l = max(l, (uint32_t)s_depth[p1 - 1] + 1u);
                        changed |= l != s_depth[i];
                    }
                    nl[j] = l;
                }
                changed = __syncthreads_or(changed);
#pragma unroll
                for (int j = 0; j < (CAP + NT - 1) / NT; j++) { const int i = t + j * NT; if (i < m) s_depth[i] = (uint16_t)nl[j]; }
                __syncthreads();
                if (!changed) break;
            }
            // ---- counting sort by depth: s_order = the steps level by level, s_lstart[d] = first slot of depth d
            for (int i = t; i < m; i += NT) atomicAdd(&s_lstart[s_depth[i] + 1], 1);
            __syncthreads();
            if (wib == 0) {
                int run = 0, nlev = 0;
                for (int base = 0; base < CAP && run < m; base += 32) {
                    const int v = s_lstart[base + lane + 1];
                    int inc = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(FULLM, inc, o); if (lane >= o) inc += u; }
                    s_lstart[base + lane + 1] = run + inc; // end of depth base+lane = start of the next depth
                    const unsigned nz = __ballot_sync(FULLM, v != 0);
                    if (nz) nlev = base + 32 - __clz(nz);
                    run += __shfl_sync(FULLM, inc, 31);
                }
                if (lane == 0) s_nlev = nlev;
            }
            __syncthreads();
            for (int i = t; i < m; i += NT) { const int d = s_depth[i]; s_order[s_lstart[d] + atomicAdd(&s_cur[d], 1)] = (uint16_t)i; }
            __syncthreads();
            // ---- the piece, level by level
            const int nlev = s_nlev;
            for (int d = 0; d < nlev; d++) {
                const int l0 = s_lstart[d], l1 = s_lstart[d + 1];
                for (int j = l0 + wib; j < l1; j += NW) {
                    const int tl = s_order[j];
                    step_core(A, (unsigned long long)(c0 + tl), s_hit[tl], s_sz[2 * tl], s_sz[2 * tl + 1], hbase, lane, s_rnd[wib], s_shift[wib], s_sbuf, &s_lock);
                }
                __syncthreads(); // the stores of this level (L2-scope) are ordered before the reads of the next one
            }
        }
    }
}

// after the batch: rebuild the rasters of the levels above dirty[o] with the tracked shifts.  32 objects per warp: the
// levels with real width by the whole warp, the tail levels by one thread per object.
__global__ void __launch_bounds__(256) k_materialize(uint32_t* __restrict__ words, const LevelMeta* __restrict__ lm, int L, int n, int* __restrict__ dirty, int* __restrict__ built_from, int grp) {
    pdl_sync(); // programmatic dependent launch: scheduled while the previous kernel drains; nothing global before this
    __shared__ uint32_t sbuf[8][2 * RB_HALF];
    const unsigned FULLM = 0xffffffffu;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int g = warp; g * grp < n; g += nwarps) { // grp (1..32) objects per warp
        int o_mine = g * grp + lane;
        int m_mine = (lane < grp && o_mine < n) ? dirty[o_mine] : L;
        unsigned todo = __ballot_sync(FULLM, m_mine < L);
        while (todo) {
            int k = __ffs(todo) - 1; todo &= todo - 1;
            int o = g * grp + k, m = __shfl_sync(FULLM, m_mine, k);
            LevelMeta mine = {};
            if (lane < L) mine = ldmeta<false>(lm + (size_t)o * L + lane);
            LevelMeta mf = shfl_meta(mine, m - 1);
            int T = grp >= 4 ? tail_level_from(lm_kh(mf), lm_kw(mf), L, m) : L; // a thread tail pays off from ~4 objects per warp
            warp_rebuild<false, true>(words, mine, L, m, lane, sbuf[threadIdx.x >> 5], T);
            __syncwarp();
        }
        if (m_mine < L) { if (grp >= 4) thread_tail_build<true>(words, lm + (size_t)o_mine * L, L, m_mine); dirty[o_mine] = L; built_from[o_mine] = m_mine - 1; }
        __syncwarp();
    }
}

// union!(moved, first.(colist) |> Iterators.flatten) (qtree_functions.jl:354-355) on the device: the distinct labels of a hit
// list in ascending order.  One CTA: endpoints stamp mark[label] with this call's epoch (no clearing between calls), then
// the labels 1..n are compacted in order (block scan per 1024-label slab).  out[0] = count, out[1..] = labels.
__global__ void __launch_bounds__(1024) k_result_labels(const Item16* __restrict__ hits, const unsigned long long* __restrict__ n_dev, int64_t n_cap,
                                                        uint32_t* __restrict__ mark, int n_obj, uint32_t epoch, int32_t* __restrict__ out, int out_cap) {
    pdl_sync();
    __shared__ int wsum[32];
    __shared__ int base;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int64_t n = min((int64_t)*n_dev, n_cap);
    for (int64_t e = t; e < 2 * n; e += 1024) { const Item16 h = hits[e >> 1]; mark[((e & 1) ? item_i2(h) : h.i1) - 1] = epoch; }
    if (t == 0) base = 0;
    __syncthreads();
    for (int o0 = 0; o0 < n_obj; o0 += 1024) {
        const int o = o0 + t;
        const bool m = o < n_obj && mark[o] == epoch;
        const unsigned bal = __ballot_sync(0xffffffffu, m);
        if (lane == 0) wsum[w] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int i = 0; i < w; i++) off += wsum[i];
        const int pos = off + __popc(bal & ((1u << lane) - 1u));
        if (m && 1 + pos < out_cap) out[1 + pos] = o + 1;
        __syncthreads();
        if (t == 0) { int sum = 0; for (int i = 0; i < 32; i++) sum += wsum[i]; base += sum; }
        __syncthreads();
    }
    if (t == 0) out[0] = base;
}

// ---- trainepoch_E! on the device (fit.jl:85-99): the end-of-round kernel of stf_fit_epoch_e.
// Round 0 ran over every label: `inds := union!(empty!(moved), first.(colist) |> Iterators.flatten)` — the distinct labels of
// the colist in order of FIRST APPEARANCE (IntSet keeps insertion order, common_datatypes.jl:227-237) — is built here
// (atomicMin of the endpoint index per label, then an ordered compaction over the endpoints), together with the round
// limit length(qtrees) ÷ length(inds).  Rounds ni >= 1 ran over `inds`: the epoch ends after the round with
// ni > 8 length(colist) (:96) or with ni == limit.  ctl = {ninds, limit, ni, reason}; reason 1 = epoch finished,
// 3 = `inds` is at or below the dispatcher's threshold (qtree_functions.jl:265-276: the native all-pairs path takes over,
// driven by the host).  *stop != 0 cancels every round enqueued after this one.
#define EPOCH_E_THRESHOLD 10
__global__ void __launch_bounds__(1024) k_epoch_e_end(unsigned long long* __restrict__ counts, int c_stop, int c_status, int c_nhits, int c_items, int c_direct, int c_visited, int c_moves, int c_nvalid,
                                                      unsigned long long* __restrict__ rec, int k_rec, int ni, const Item16* __restrict__ hits, int64_t hits_cap,
                                                      int* __restrict__ firstpos, int n_obj, int32_t* __restrict__ inds, int* __restrict__ ctl) {
    pdl_sync();
    __shared__ int wsum[32];
    __shared__ int base, s_go;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) {
        int go = 1;
        if (counts[c_stop] != 0ull) go = 0;                                        // cancelled round: no record
        else if ((int)counts[c_status] != 0) { counts[c_stop] = 2ull; go = 0; }    // this round raised a bit (its steps did not run)
        if (go) {
            unsigned long long* r = rec + (size_t)4 * (k_rec + 1);
            r[0] = counts[c_nhits]; r[1] = counts[c_items] + counts[c_direct]; r[2] = counts[c_visited]; r[3] = counts[c_moves];
            rec[0] = (unsigned long long)(k_rec + 1); rec[1] = counts[c_nvalid];
        }
        s_go = go; base = 0;
    }
    __syncthreads();
    if (!s_go) return;
    const int64_t nh = min((int64_t)counts[c_nhits], hits_cap);
    if (ni > 0) { // a round over `inds`
        if (t == 0) { ctl[2] = ni; if ((int64_t)ni > 8 * nh || ni >= ctl[1]) { ctl[3] = 1; counts[c_stop] = 1ull; } }
        return;
    }
    if (nh == 0) { if (t == 0) { ctl[0] = 0; ctl[1] = 0; ctl[2] = 0; ctl[3] = 1; counts[c_stop] = 1ull; } return; } // nc == 0: return nc (:89)
    const int64_t ne = 2 * nh;
    for (int64_t e = t; e < ne; e += 1024) { const Item16 h = hits[e >> 1]; firstpos[((e & 1) ? item_i2(h) : h.i1) - 1] = 0x7fffffff; }
    __syncthreads();
    for (int64_t e = t; e < ne; e += 1024) { const Item16 h = hits[e >> 1]; atomicMin(firstpos + (((e & 1) ? item_i2(h) : h.i1) - 1), (int)e); }
    __syncthreads();
    for (int64_t e0 = 0; e0 < ne; e0 += 1024) {
        const int64_t e = e0 + t;
        int label = 0; bool first = false;
        if (e < ne) { const Item16 h = hits[e >> 1]; label = (e & 1) ? item_i2(h) : h.i1; first = firstpos[label - 1] == (int)e; }
        const unsigned bal = __ballot_sync(0xffffffffu, first);
        if (lane == 0) wsum[w] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int i = 0; i < w; i++) off += wsum[i];
        if (first) inds[off + __popc(bal & ((1u << lane) - 1u))] = label;
        __syncthreads();
        if (t == 0) { int sum = 0; for (int i = 0; i < 32; i++) sum += wsum[i]; base += sum; }
        __syncthreads();
    }
    if (t == 0) {
        const int ninds = base;
        ctl[0] = ninds; ctl[1] = n_obj / ninds; ctl[2] = 0; ctl[3] = 0;
        if (ctl[1] < 1) { ctl[3] = 1; counts[c_stop] = 1ull; }                    // for ni in 1:0 — no inner round
        else if (ninds <= EPOCH_E_THRESHOLD) { ctl[3] = 3; counts[c_stop] = 3ull; } // the native path takes over (host)
    }
}
