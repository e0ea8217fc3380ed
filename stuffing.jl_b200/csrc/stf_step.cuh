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
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * n) return;
    int64_t k = e >> 1; int s = (int)(e & 1);
    int o = (s ? item_i2(hits[k]) : hits[k].i1) - 1;
    keys[e] = om[o].is_mask ? STF_KEY_INVALID : (uint64_t)(uint32_t)o;
    vals[e] = (uint32_t)e;
}
__global__ void k_step_prev(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const unsigned long long* __restrict__ m_dev, int64_t m_cap, uint32_t* __restrict__ prev) {
    int64_t m = m_dev ? min((int64_t)*m_dev, m_cap) : m_cap;
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    uint64_t key = keys[p];
    uint32_t e = vals[p];
    uint32_t pv = 0;
    if (key != STF_KEY_INVALID && p > 0 && keys[p - 1] == key) pv = (vals[p - 1] >> 1) + 1;
    prev[e] = pv;
}

__device__ __forceinline__ int intlog2_dev(double x) { return (int)((__double_as_longlong(x) >> 52) - 1023); } // fit_helper.jl:74-78

struct StepArgs {
    uint32_t* words; LevelMeta* lm; const ObjMeta* om; int L;
    const Item16* hits; int64_t n;
    const uint32_t* prev; uint32_t* done; unsigned long long* ticket;
    double* vel; uint8_t* has_vel; uint8_t* moved_flag;
    OptState opt; uint64_t seed, iter;
    unsigned long long* counts; // [4] number of moves
    unsigned long long* dbg_t;  // optional timeline (4 x u64 per step) when dbg & 16
    int dbg;                    // STF_DEBUG_STEP experiments: 1 no wait, 2 no rebuild, 4 no fence (breaks parity)
};

// optimiser(t, Δ)  fit_helper.jl:19-35, fit.jl:25 — executed by one lane; (has, v0, v1) were prefetched
__device__ __forceinline__ void opt_apply_dev(const StepArgs& A, int o, double& d0, double& d1, int has, double v0, double v1) {
    if (A.opt.kind == 0) { d0 = d0 / 6.0; d1 = d1 / 6.0; return; }
    if (A.opt.kind == 1) { d0 = __dmul_rn(A.opt.eta, d0); d1 = __dmul_rn(A.opt.eta, d1); return; }
    if (!has) { v0 = d0; v1 = d1; __stcg(A.has_vel + o, (uint8_t)1); } // v = get!(velocity, x, Δ)
    double omr = __dsub_rn(1.0, A.opt.rho);
    v0 = __dadd_rn(__dmul_rn(A.opt.rho, v0), __dmul_rn(omr, d0)); // no FMA contraction: Julia does not fuse
    v1 = __dadd_rn(__dmul_rn(A.opt.rho, v1), __dmul_rn(omr, d1));
    __stcg(A.vel + 2 * (size_t)o, v0); __stcg(A.vel + 2 * (size_t)o + 1, v1);
    d0 = __dmul_rn(A.opt.eta, v0); d1 = __dmul_rn(A.opt.eta, v1);
}
// move!  fit.jl:11-23 — decides (level, s1, s2); executed by one lane.  r_jit, r_d1, r_d2: the three draws of this
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
// shift!(t, lv, s1, s2)  qtrees.jl:267-275 by one warp.  `mine`: lane i holds the meta of level i+1 of object o.
// Levels <= lv move by s*2^(lv-i); levels > lv are rebuilt; all changed metas are stored once.
__device__ __forceinline__ void warp_shift_rebuild(const StepArgs& A, int o, LevelMeta mine, int3 mv, int lane, uint32_t* sbuf) {
    int lv = mv.x;
    if (lane < lv) { int f = 1 << (lv - 1 - lane); mine.rs += mv.y * f; mine.cs += mv.z * f; }
    mine = warp_rebuild<true>(A.words, mine, A.L, lv, lane, sbuf);
    if (lane < A.L) stmeta<true>(A.lm + (size_t)o * A.L + lane, mine);
    if (lane == 0) { A.moved_flag[o] = 1; atomicAdd(A.counts + 4, 1ull); }
}

// Per step, dependent L2 round trips: (1) wait flags, (2) both objects' LevelMetas of every level (one per
// lane: lanes 0-15 tree 1, 16-31 tree 2) + velocities, (3) the 2x8 gradient nodes (one per lane), (4) the
// child words of the first rebuilt level; everything above is rebuilt through shared memory.
__global__ void __launch_bounds__(128) k_batchsteps(StepArgs A) {
    __shared__ uint32_t sbuf[4][2 * RB_HALF];
    int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, half = lane >> 4, hl = lane & 15;
    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(A.ticket, 1ull);
        k = __shfl_sync(0xffffffffu, k, 0);
        if ((int64_t)k >= A.n) return;
        unsigned long long t_claim = 0, t_ready = 0, t_decided = 0;
        if (A.dbg & 16) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_claim));
        Item16 h = A.hits[k];
        int o1 = h.i1 - 1, o2 = item_i2(h) - 1;
        int l = item_l(h), a = h.a, b = h.b;
        ObjMeta om1 = A.om[o1], om2 = A.om[o2];
        if (lane < 2 && !(A.dbg & 1)) { // wait for the previous step of each movable object
            uint32_t pv = A.prev[2 * k + lane];
            if (pv) { const volatile uint32_t* f = A.done + (pv - 1); if (A.dbg & 8) { while (*f == 0) __nanosleep(20); } else { while (*f == 0) { } } }
        }
        __syncwarp();
        __threadfence();
        if (A.dbg & 16) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ready));
        // (2) metas of all levels of both trees + optimiser state
        int oh = half ? o2 : o1;
        LevelMeta mine = {};
        if (hl < A.L) mine = ldmeta<true>(A.lm + (size_t)oh * A.L + hl);
        int has = 0; double v0 = 0, v1 = 0;
        if (hl == 0 && A.opt.kind == 2) {
            has = __ldcg(A.has_vel + oh);
            double2 vv = __ldcg(reinterpret_cast<const double2*>(A.vel + 2 * (size_t)oh));
            v0 = vv.x; v1 = vv.y;
        }
        // (3) grad2d (fit_helper.jl:51-66): gh = sum dr*decode2, gw = sum dc*decode2 over the 8 neighbours
        LevelMeta q = shfl_meta(mine, half * 16 + (l - 1));
        int gh = 0, gw = 0;
        if (hl < 8) {
            int idx = hl < 4 ? hl : hl + 1; int dr = idx / 3 - 1, dc = idx % 3 - 1;
            int v = decode2(node_code<true>(A.words, q, a + dr, b + dc));
            gh = v * dr; gw = v * dc;
        }
#pragma unroll
        for (int s = 1; s < 8; s <<= 1) { gh += __shfl_xor_sync(0xffffffffu, gh, s); gw += __shfl_xor_sync(0xffffffffu, gw, s); }
        int g1x = __shfl_sync(0xffffffffu, gh, 0), g1y = __shfl_sync(0xffffffffu, gw, 0);
        int g2x = __shfl_sync(0xffffffffu, gh, 16), g2y = __shfl_sync(0xffffffffu, gw, 16);
        int has2 = __shfl_sync(0xffffffffu, has, 16); double v20 = __shfl_sync(0xffffffffu, v0, 16), v21 = __shfl_sync(0xffffffffu, v1, 16);
        // the 8 draws of this step, one per lane (stuffing_b200_rng.h slots 0..7), gathered by lane 0
        uint64_t rnd = stf_rand_u64(A.seed, A.iter, k, (uint32_t)(lane & 7));
        uint64_t R[8];
#pragma unroll
        for (int j = 0; j < 8; j++) R[j] = __shfl_sync(0xffffffffu, rnd, j);
        bool m1 = om1.is_mask, m2 = om2.is_mask;
        int3 mvA = make_int3(0, 0, 0), mvB = make_int3(0, 0, 0); int oA = -1, oB = -1;
        if (lane == 0) { // the scalar arithmetic of step!/step_mask!
            double ll = (double)(1ll << (l - 1));
            double d1x = ll * g1x, d1y = ll * g1y, d2x = ll * g2x, d2y = ll * g2y;
            if (m1 || m2) { // step_mask!  fit.jl:49-56 (i1 == 1 first, then i2 == 1: fit.jl:60-63)
                int ot = m1 ? o2 : o1;
                double dx = m1 ? (d2x - d1x) / 2 : (d1x - d2x) / 2, dy = m1 ? (d2y - d1y) / 2 : (d1y - d2y) / 2;
                if (m1) opt_apply_dev(A, ot, dx, dy, has2, v20, v21); else opt_apply_dev(A, ot, dx, dy, has, v0, v1);
                mvA = move_decide(A, dx, dy, R[2], R[3], R[4]); oA = ot;
            } else { // step!  fit.jl:25-48
                double s1 = (double)(om1.kh1 + om1.kw1), s2 = (double)(om2.kh1 + om2.kw1);
                double ks1 = s1 * s1 * s1, ks2 = s2 * s2 * s2;
                bool move1 = u64_to_unit(R[0]) < ks2 / ks1;
                bool move2 = u64_to_unit(R[1]) < ks1 / ks2;
                int slot = 2;
                if (move1) {
                    double dx = d1x, dy = d1y;
                    if (!move2) { dx -= d2x; dy -= d2y; }
                    opt_apply_dev(A, o1, dx, dy, has, v0, v1);
                    mvA = move_decide(A, dx, dy, R[2], R[3], R[4]); oA = o1; slot += 3;
                }
                if (move2) {
                    double dx = d2x, dy = d2y;
                    if (!move1) { dx -= d1x; dy -= d1y; }
                    opt_apply_dev(A, o2, dx, dy, has2, v20, v21);
                    int3 mv = slot == 2 ? move_decide(A, dx, dy, R[2], R[3], R[4]) : move_decide(A, dx, dy, R[5], R[6], R[7]);
                    if (oA < 0) { mvA = mv; oA = o2; } else { mvB = mv; oB = o2; }
                }
            }
        }
        oA = __shfl_sync(0xffffffffu, oA, 0); oB = __shfl_sync(0xffffffffu, oB, 0);
        mvA.x = __shfl_sync(0xffffffffu, mvA.x, 0); mvA.y = __shfl_sync(0xffffffffu, mvA.y, 0); mvA.z = __shfl_sync(0xffffffffu, mvA.z, 0);
        mvB.x = __shfl_sync(0xffffffffu, mvB.x, 0); mvB.y = __shfl_sync(0xffffffffu, mvB.y, 0); mvB.z = __shfl_sync(0xffffffffu, mvB.z, 0);
        long long c_decided = 0;
        if (A.dbg & 16) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_decided)); c_decided = clock64(); }
        // metas re-laid so that lane i holds level i+1 of the tree to rebuild
        LevelMeta t1 = shfl_meta(mine, lane & 15), t2 = shfl_meta(mine, 16 + (lane & 15));
        if (lane >= 16) { t1 = LevelMeta(); t2 = LevelMeta(); }
        if (A.dbg & 2) { oA = -1; oB = -1; }
        if (oA >= 0) { warp_shift_rebuild(A, oA, oA == o1 ? t1 : t2, mvA, lane, sbuf[wib]); __syncwarp(); }
        if (oB >= 0) { warp_shift_rebuild(A, oB, oB == o1 ? t1 : t2, mvB, lane, sbuf[wib]); __syncwarp(); }
        long long c_rebuilt = (A.dbg & 16) ? clock64() : 0;
        if (!(A.dbg & 4)) __threadfence();
        __syncwarp();
        if (lane == 0) {
            if (A.dbg & 16) { unsigned long long t_end; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end)); A.dbg_t[4 * k] = ((unsigned long long)(c_rebuilt - c_decided) << 32) | (unsigned long long)(clock64() - c_rebuilt); A.dbg_t[4 * k + 1] = t_ready; A.dbg_t[4 * k + 2] = t_decided; A.dbg_t[4 * k + 3] = t_end; }
            *((volatile uint32_t*)(A.done + k)) = 1u;
        }
    }
}
