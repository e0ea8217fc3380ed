/* stuffing_b200_rng.h — the canonical-mode random draws of the fit step.
 *
 * The reference draws from Julia's task-local RNG (rand() in step!/move!,
 * /root/reference/src/fit.jl:12-16,34-35).  That stream cannot be reproduced
 * outside Julia, so the boundary defines a counter-based generator instead:
 * every draw is a pure function of (seed, iteration, hit index, slot).  The CPU
 * oracle and the CUDA kernels include this same header, which is what makes
 * fitted positions comparable bit for bit.
 *
 * Slots per hit (fixed, whether or not the draw is consumed):
 *   0  move1 test          rand() < ks2/ks1          fit.jl:34
 *   1  move2 test          rand() < ks1/ks2          fit.jl:35
 *   2  first moved tree:   jitter test  rand() < 0.1 fit.jl:12
 *   3  first moved tree:   jitter diagonal           fit.jl:13
 *   4  first moved tree:   "avoid rest" diagonal     fit.jl:16
 *   5,6,7  the same three draws for the second moved tree
 * step_mask! (fit.jl:49-56) moves one tree and uses slots 2..4.
 */
#ifndef STUFFING_B200_RNG_H
#define STUFFING_B200_RNG_H
#include <stdint.h>

#if defined(__CUDACC__)
#define STF_HD __host__ __device__ __forceinline__
#else
#define STF_HD static inline
#endif

STF_HD uint64_t stf_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* 64 random bits for (seed, iteration, hit index k, slot 0..7) */
STF_HD uint64_t stf_rand_u64(uint64_t seed, uint64_t iter, uint64_t k, uint32_t slot) {
    uint64_t h = stf_mix64(seed ^ 0x5354554646494E47ull); /* "STUFFING" */
    h = stf_mix64(h + iter);
    h = stf_mix64(h + (k << 3) + (uint64_t)slot);
    return h;
}

/* uniform double in [0,1) with 53 random bits, like Julia's rand(Float64) */
STF_HD double stf_rand_f64(uint64_t seed, uint64_t iter, uint64_t k, uint32_t slot) {
    return (double)(stf_rand_u64(seed, iter, k, slot) >> 11) * (1.0 / 9007199254740992.0);
}

/* index 0..3 into ((1,-1),(-1,1),(-1,-1),(1,1)) — fit.jl:13,16 */
STF_HD int stf_rand_diag(uint64_t seed, uint64_t iter, uint64_t k, uint32_t slot) {
    return (int)(stf_rand_u64(seed, iter, k, slot) >> 62);
}

#endif
