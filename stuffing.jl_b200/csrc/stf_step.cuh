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
__global__ void k_step_prev(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t m, uint32_t* __restrict__ prev) {
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
};

// optimiser(t, Δ)  fit_helper.jl:19-35, fit.jl:25  — executed by one lane
__device__ __forceinline__ void opt_apply_dev(const StepArgs& A, int o, double& d0, double& d1) {
    if (A.opt.kind == 0) { d0 = d0 / 6.0; d1 = d1 / 6.0; return; }
    if (A.opt.kind == 1) { d0 = __dmul_rn(A.opt.eta, d0); d1 = __dmul_rn(A.opt.eta, d1); return; }
    double v0, v1;
    if (!__ldcg(A.has_vel + o)) { v0 = d0; v1 = d1; __stcg(A.has_vel + o, (uint8_t)1); }
    else { v0 = __ldcg(A.vel + 2 * (size_t)o); v1 = __ldcg(A.vel + 2 * (size_t)o + 1); }
    double omr = __dsub_rn(1.0, A.opt.rho);
    v0 = __dadd_rn(__dmul_rn(A.opt.rho, v0), __dmul_rn(omr, d0)); // no FMA contraction: Julia does not fuse
    v1 = __dadd_rn(__dmul_rn(A.opt.rho, v1), __dmul_rn(omr, d1));
    __stcg(A.vel + 2 * (size_t)o, v0); __stcg(A.vel + 2 * (size_t)o + 1, v1);
    d0 = __dmul_rn(A.opt.eta, v0); d1 = __dmul_rn(A.opt.eta, v1);
}
// move!  fit.jl:11-23 — decides (level, s1, s2); executed by one lane
__device__ __forceinline__ int3 move_decide(const StepArgs& A, double d0, double d1, uint64_t k, uint32_t slot0) {
    const double DG[4][2] = { { 1., -1. }, { -1., 1. }, { -1., -1. }, { 1., 1. } };
    if (stf_rand_f64(A.seed, A.iter, k, slot0) < 0.1) { int j = stf_rand_diag(A.seed, A.iter, k, slot0 + 1); d0 = __dadd_rn(d0, DG[j][0]); d1 = __dadd_rn(d1, DG[j][1]); }
    if (-1 < d0 && d0 < 1 && -1 < d1 && d1 < 1) { int j = stf_rand_diag(A.seed, A.iter, k, slot0 + 2); d0 = DG[j][0]; d1 = DG[j][1]; }
    double dm = fmax(fabs(d0), fabs(d1));
    int u = intlog2_dev(dm);
    long long p = 1ll << u;
    int s1 = (int)((long long)trunc(d0) / p), s2 = (int)((long long)trunc(d1) / p);
    return make_int3(min(A.L, 1 + u), s1, s2);
}
// shift!(t, lv, s1, s2)  qtrees.jl:267-275 by one warp: levels <= lv move, levels > lv are rebuilt
__device__ __forceinline__ void warp_shift_rebuild(const StepArgs& A, int o, int3 mv, int lane) {
    int lv = mv.x;
    if (lane < lv) { // level i = lane+1 moves by s * 2^(lv-i)
        LevelMeta* p = A.lm + (size_t)o * A.L + lane;
        LevelMeta m = ldmeta<true>(p);
        int f = 1 << (lv - 1 - lane);
        uint4 v = make_uint4(m.off, (uint32_t)(m.rs + mv.y * f), (uint32_t)(m.cs + mv.z * f), m.dims);
        __stcg(reinterpret_cast<uint4*>(p), v);
    }
    __syncwarp();
    rebuild_levels_warp<true>(A.words, A.lm, o, A.L, lv, lane);
    if (lane == 0) { A.moved_flag[o] = 1; atomicAdd(A.counts + 4, 1ull); }
}

__global__ void __launch_bounds__(128) k_batchsteps(StepArgs A) {
    int lane = threadIdx.x & 31;
    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(A.ticket, 1ull);
        k = __shfl_sync(0xffffffffu, k, 0);
        if ((int64_t)k >= A.n) return;
        Item16 h = A.hits[k];
        int o1 = h.i1 - 1, o2 = item_i2(h) - 1;
        int l = item_l(h), a = h.a, b = h.b;
        if (lane < 2) { // wait for the previous step of each movable object
            uint32_t pv = A.prev[2 * k + lane];
            if (pv) { const volatile uint32_t* f = A.done + (pv - 1); while (*f == 0) __nanosleep(40); }
        }
        __syncwarp();
        __threadfence();
        bool m1 = A.om[o1].is_mask, m2 = A.om[o2].is_mask;
        // lane 0 decides the moves (the arithmetic of step!/step_mask! is a few dozen scalar operations)
        int3 mvA = make_int3(0, 0, 0), mvB = make_int3(0, 0, 0); int oA = -1, oB = -1;
        if (lane == 0) {
            LevelMeta q1 = ldmeta<true>(A.lm + (size_t)o1 * A.L + (l - 1)), q2 = ldmeta<true>(A.lm + (size_t)o2 * A.L + (l - 1));
            int2 g1 = grad2d_dev<true>(A.words, q1, a, b), g2 = grad2d_dev<true>(A.words, q2, a, b);
            double ll = (double)(1ll << (l - 1));
            double d1x = ll * g1.x, d1y = ll * g1.y, d2x = ll * g2.x, d2y = ll * g2.y;
            if (m1 || m2) { // step_mask!  fit.jl:49-56 (i1 == 1 first, then i2 == 1: fit.jl:60-63)
                int ot = m1 ? o2 : o1;
                double dx = m1 ? (d2x - d1x) / 2 : (d1x - d2x) / 2, dy = m1 ? (d2y - d1y) / 2 : (d1y - d2y) / 2;
                opt_apply_dev(A, ot, dx, dy);
                mvA = move_decide(A, dx, dy, k, 2); oA = ot;
            } else { // step!  fit.jl:25-48
                double s1 = (double)(A.om[o1].kh1 + A.om[o1].kw1), s2 = (double)(A.om[o2].kh1 + A.om[o2].kw1);
                double ks1 = s1 * s1 * s1, ks2 = s2 * s2 * s2;
                bool move1 = stf_rand_f64(A.seed, A.iter, k, 0) < ks2 / ks1;
                bool move2 = stf_rand_f64(A.seed, A.iter, k, 1) < ks1 / ks2;
                uint32_t slot = 2;
                if (move1) {
                    double dx = d1x, dy = d1y;
                    if (!move2) { dx -= d2x; dy -= d2y; }
                    opt_apply_dev(A, o1, dx, dy);
                    mvA = move_decide(A, dx, dy, k, slot); oA = o1; slot += 3;
                }
                if (move2) {
                    double dx = d2x, dy = d2y;
                    if (!move1) { dx -= d1x; dy -= d1y; }
                    opt_apply_dev(A, o2, dx, dy);
                    int3 mv = move_decide(A, dx, dy, k, slot);
                    if (oA < 0) { mvA = mv; oA = o2; } else { mvB = mv; oB = o2; }
                }
            }
        }
        oA = __shfl_sync(0xffffffffu, oA, 0); oB = __shfl_sync(0xffffffffu, oB, 0);
        mvA.x = __shfl_sync(0xffffffffu, mvA.x, 0); mvA.y = __shfl_sync(0xffffffffu, mvA.y, 0); mvA.z = __shfl_sync(0xffffffffu, mvA.z, 0);
        mvB.x = __shfl_sync(0xffffffffu, mvB.x, 0); mvB.y = __shfl_sync(0xffffffffu, mvB.y, 0); mvB.z = __shfl_sync(0xffffffffu, mvB.z, 0);
        if (oA >= 0) warp_shift_rebuild(A, oA, mvA, lane);
        if (oB >= 0) warp_shift_rebuild(A, oB, mvB, lane);
        __syncwarp();
        __threadfence();
        if (lane == 0) { *((volatile uint32_t*)(A.done + k)) = 1u; }
    }
}
