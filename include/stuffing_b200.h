/* stuffing_b200.h — C ABI of the B200-native collision-and-fit engine.
 *
 * Drop-in boundary for the hot path of guo-yong-zhi/Stuffing.jl v0.10.4.  Every entry point
 * names the reference interface it replaces (file:line under /root/reference/src/).  The
 * reference is pure Julia; the binding a maintainer adds is a `ccall` per function
 * (julia/StuffingB200.jl).  No torch types, no C++ types: plain pointers,
 * sizes and int32 status codes.
 *
 * Conventions
 *  - labels and (l,a,b) are 1-based, exactly as the reference prints them;
 *  - matrices are column-major (row index fastest), like Julia's Matrix{UInt8};
 *  - node codes: EMPTY=1, FULL=2, MIX=3 (qtrees.jl:56);
 *  - a context owns device copies of one vector of ShiftedQTrees ("objects") that share a
 *    canvas; the device is authoritative for shifts and levels >= 2 after upload;
 *  - every call is synchronous with respect to its outputs; calls on one context must be
 *    serialised by the caller;
 *  - result lists use the two-call pattern: if `cap` is too small the call returns
 *    STF_E_CAPACITY, stores the needed size in *count, and keeps the result on the device for
 *    stf_fetch_result();
 *  - there is no CPU fallback: without a CUDA device stf_create fails with STF_E_CUDA.
 */
#ifndef STUFFING_B200_H
#define STUFFING_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define STF_EMPTY 1
#define STF_FULL 2
#define STF_MIX 3

#define STF_OK 0
#define STF_E_CUDA (-1)      /* CUDA runtime error (text in stf_last_error) */
#define STF_E_ARG (-2)       /* bad argument */
#define STF_E_CAPACITY (-3)  /* output buffer too small; *count holds the needed size */
#define STF_E_BADCODE (-4)   /* a node code outside 1..3 was uploaded */
#define STF_E_RANGE (-5)     /* a spatial cell coordinate left the representable window */
#define STF_E_STATE (-6)     /* call not valid in this state (nothing uploaded, ...) */

typedef struct stf_ctx stf_ctx;

/* CoItem = Pair{Tuple{Int,Int},Index}: (i1,i2) => (l,a,b)   qtree_functions.jl:52 */
typedef struct { int32_t i1, i2, l, a, b; } stf_item;
/* one spatial-tree entry: `label` sits in cell (l,a,b)       qtrees.jl:411-415,501-524 */
typedef struct { int32_t l, a, b, label; } stf_cellrec;

/* optimiser kinds (fit_helper.jl:13-38, fit.jl:25) */
#define STF_OPT_DEFAULT 0 /* (t, Δ) -> Δ ./ 6 */
#define STF_OPT_SGD 1
#define STF_OPT_MOMENTUM 2

/* counters of the last many-body call */
typedef struct {
    int64_t records;  /* located (cell,label) records          */
    int64_t items;    /* narrow-phase items (the reference's itemlist) */
    int64_t direct;   /* default-FULL direct hits (qtree_functions.jl:212-218) */
    int64_t hits;     /* results before unique                  */
    int64_t overflow; /* items that needed the large-frontier path */
    int64_t moved;    /* objects moved by the last batchsteps   */
    int64_t visited;  /* child nodes tested by the narrow phase */
    int64_t reserved;
} stf_counters_t;

/* ---- lifetime ---- */
int32_t stf_create(int32_t device, stf_ctx** out);
void stf_destroy(stf_ctx* ctx);
const char* stf_last_error(const stf_ctx* ctx);
/* the context's CUDA stream (cudaStream_t) so that callers can time with events on it */
void* stf_stream(stf_ctx* ctx);

/* ---- construction: qtree / maskqtree / qtrees (Stuffing.jl:16-63) + ShiftedQTree constructors
 * (qtrees.jl:176-209) + buildqtree! (qtrees.jl:221-237).
 * pics[i] points at the kh[i] x kw[i] level-1 payload (codes 1..3), column-major with leading
 * dimension pitch[i] (Julia: pointer(kernel, 2,2) with pitch kh+3, or the pic itself).
 * rshift/cshift: level-1 shifts (maskqtree centres the mask; qtree uses 0).  dflt: EMPTY for
 * objects, FULL for the mask.  canvas: power of two >= 2.  is_mask (nullable): which objects
 * step_index! treats as the mask (default: label 1 only, fit.jl:58-67 — pass all zeros for
 * scenes without a mask).  scene (nullable): scene id per object for batches of independent
 * scenes that share a canvas size; objects of different scenes never interact.
 * All levels are built on the device before the call returns. */
int32_t stf_upload_objects(stf_ctx* ctx, int32_t n, const uint8_t* const* pics, const int32_t* kh,
                           const int32_t* kw, const int32_t* pitch, const int32_t* rshift,
                           const int32_t* cshift, const uint8_t* dflt, int32_t canvas,
                           const uint8_t* is_mask, const int32_t* scene);
/* same, with all level-1 payloads packed back to back (kh[i]*kw[i] bytes each, pitch = kh[i]) in one
 * buffer that is ALREADY IN DEVICE MEMORY (bench "inputs resident in HBM" leg) or in host memory */
int32_t stf_upload_packed(stf_ctx* ctx, int32_t n, const uint8_t* payload, int32_t payload_on_device,
                          const int32_t* kh, const int32_t* kw, const int32_t* rshift,
                          const int32_t* cshift, const uint8_t* dflt, int32_t canvas,
                          const uint8_t* is_mask, const int32_t* scene);
int32_t stf_num_objects(const stf_ctx* ctx);
int32_t stf_num_levels(const stf_ctx* ctx); /* length(t)  qtrees.jl:217 */

/* ---- tree state ---- */
/* out = {kh, kw, rshift, cshift, canvas size} of level l: kernelsize / getshift / size  qtrees.jl:152,288-289 */
int32_t stf_get_level_info(stf_ctx* ctx, int32_t label, int32_t l, int32_t out[5]);
/* kh*kw payload bytes of level l, column-major: t[l].kernel[2:end-2,2:end-2] */
int32_t stf_download_level(stf_ctx* ctx, int32_t label, int32_t l, uint8_t* out);
/* every level of the listed trees (labels == NULL: all) in one call: for each tree, levels 1..L back to back, each
 * kh_l*kw_l bytes column-major (the payloads of t[1..L]; kh_{l+1} = kh_l ÷ 2 + 1, qtrees.jl:185).  *count = total bytes
 * (two-call pattern with cap).  One kernel launch instead of one per (tree, level). */
int32_t stf_download_trees(stf_ctx* ctx, int32_t n, const int32_t* labels, uint8_t* out, int64_t cap, int64_t* count);
/* t[l, a, b] for n coordinates: PaddedMat getindex, `default` outside the kernel  qtrees.jl:156-162 */
int32_t stf_read_nodes(stf_ctx* ctx, int32_t n, const int32_t* labels, const int32_t* l,
                       const int32_t* a, const int32_t* b, uint8_t* out);
int32_t stf_get_shifts(stf_ctx* ctx, int32_t n, const int32_t* labels, int32_t level, int32_t* rs, int32_t* cs);
/* mode 0: setshift!(t, level, rs, cs) (qtrees.jl:277-285); mode 1: shift!(t, level, rs, cs) (:267-275).
 * Levels <= level move, levels > level are rebuilt.  labels == NULL: objects 1..n.
 * mode | STF_SHIFT_DEVICE_ARRAYS: rs/cs are device pointers (inputs already resident in HBM).
 * mode | STF_SHIFT_DEVICE_ARRAYS | STF_SHIFT_DEVICE_LABELS: `labels` is a device pointer too (a resident list: no host pass
 * over it, no copy).  The kernel validates it: an out-of-range label or a mask-sized object in the list is dropped and
 * reported as STF_E_ARG by the next synchronising call. */
#define STF_SHIFT_DEVICE_ARRAYS 2
#define STF_SHIFT_DEVICE_LABELS 4
int32_t stf_set_shifts(stf_ctx* ctx, int32_t n, const int32_t* labels, int32_t level,
                       const int32_t* rs, const int32_t* cs, int32_t mode);
/* setrshift!(t[l], v) / setcshift!(t[l], v) on one PaddedMat, no rebuild  qtrees.jl:134-135 */
int32_t stf_set_level_shift(stf_ctx* ctx, int32_t label, int32_t l, int32_t rs, int32_t cs);
/* buildqtree!(t, layer) for the given labels (NULL: all)  qtrees.jl:221-237 */
int32_t stf_build(stf_ctx* ctx, int32_t n, const int32_t* labels, int32_t layer);

/* ---- narrow phase ---- */
/* collision(Q1,Q2)  qtree_functions.jl:43-51.  out = (l,a,b); l < 0: no collision */
int32_t stf_collision(stf_ctx* ctx, int32_t i1, int32_t i2, int32_t out[3]);
/* _collision_randbfs from explicit start nodes for n pairs, canonical child order; returns EVERY
 * result (hit or (-l,a,b) miss) in item order  qtree_functions.jl:21-42 */
int32_t stf_collision_bfs(stf_ctx* ctx, int64_t n, const stf_item* items, stf_item* out);
/* collision_dfs(Q1,Q2,i)  qtree_functions.jl:2-20, n pairs */
int32_t stf_collision_dfs(stf_ctx* ctx, int64_t n, const stf_item* items, stf_item* out);
/* _totalcollisions_native(qtrees, coitems, colist)  qtree_functions.jl:89-110: hits only */
int32_t stf_collide_items(stf_ctx* ctx, int64_t n, const stf_item* items, stf_item* out, int64_t cap, int64_t* count);

/* the pool test of the pairwise trainers: filttrain!(qtrees, inpool, outpool, nearlevel2)  fit.jl:235-266.  `pairs` =
 * n_pairs x 2 labels, or NULL = all pairs i < j of the objects, generated ON THE DEVICE in the reference's order (fit.jl:279;
 * 5e7 pairs at 10 k objects never cross the boundary).  Every pair is tested from the root (L,1,1), no bounds pre-check;
 * the pairs whose result level is >= nearlevel come back as (i1,i2) => (l,a,b) in pool order (l > 0: a collision to step;
 * nearlevel <= l < 0: a near miss, -l = level of the deepest node still MIX in both trees).  Two-call pattern on cap. */
int32_t stf_filter_pairs(stf_ctx* ctx, int64_t n_pairs, const int32_t* pairs, int32_t nearlevel, stf_item* out, int64_t cap, int64_t* count);

/* ---- many-body collision ---- */
/* totalcollisions_native(qtrees, labels)  qtree_functions.jl:111-131 (labels == NULL: 1..n) */
int32_t stf_totalcollisions_native(stf_ctx* ctx, int32_t nl, const int32_t* labels, stf_item* out, int64_t cap, int64_t* count);
/* locate!(qtrees[, labels], HashSpacialQTree)  qtree_functions.jl:159-201: records sorted by cell
 * (level, a, b), labels of one cell in `labels` order */
int32_t stf_locate(stf_ctx* ctx, int32_t nl, const int32_t* labels, stf_cellrec* out, int64_t cap, int64_t* count);
/* totalcollisions_spacial(qtrees[, labels]; unique)  qtree_functions.jl:225-263.
 * unique != 0: sorted by ((i1,i2),(l,a,b)), first entry per unordered pair (:252).
 * unique == 0: all hits, same sort order (the reference's order is arbitrary). */
int32_t stf_totalcollisions_spacial(stf_ctx* ctx, int32_t nl, const int32_t* labels, int32_t unique,
                                    stf_item* out, int64_t cap, int64_t* count);
/* the narrow-phase item list of the last spacial/partial call (the reference's `itemlist` keyword buffer,
 * qtree_functions.jl:225-249), unordered.  By default the candidate pairs are tested where they are generated and never
 * stored (only counted: stf_counters_t.items); stf_set_keep_items(ctx, 1) makes the following many-body calls
 * materialise the list so that it can be fetched. */
int32_t stf_set_keep_items(stf_ctx* ctx, int32_t on);
int32_t stf_fetch_items(stf_ctx* ctx, stf_item* out, int64_t cap, int64_t* count);
int32_t stf_fetch_result(stf_ctx* ctx, stf_item* out, int64_t cap, int64_t* count);

/* LinkedSpacialQTree state kept in the context (qtrees.jl:424-552; linked_spacial_qtree):
 * locate!(qtrees, labels, linkedtree) relocates `labels` (cells outside the root are dropped, :506) */
int32_t stf_linked_reset(stf_ctx* ctx);
int32_t stf_linked_locate(stf_ctx* ctx, int32_t nl, const int32_t* labels);
int32_t stf_linked_clear(stf_ctx* ctx, int32_t nl, const int32_t* labels); /* clear!(t,label) :530-538 */
int32_t stf_linked_collect(stf_ctx* ctx, stf_cellrec* out, int64_t cap, int64_t* count); /* collect_tree */
/* partialcollisions(qtrees, linkedtree, labels; unique)  qtree_functions.jl:277-330 */
int32_t stf_partialcollisions(stf_ctx* ctx, int32_t nl, const int32_t* labels, int32_t unique,
                              stf_item* out, int64_t cap, int64_t* count);

/* ---- fit step ---- */
/* grad2d(t, l, a, b) -> (gh, gw) for n queries  fit_helper.jl:51-66; out = 2n int32 */
int32_t stf_grad2d(stf_ctx* ctx, int32_t n, const int32_t* labels, const int32_t* l, const int32_t* a,
                   const int32_t* b, int32_t* out);
int32_t stf_set_optimiser(stf_ctx* ctx, int32_t kind, double eta, double rho); /* Momentum/SGD; eta! */
int32_t stf_reset_velocity(stf_ctx* ctx, int32_t n, const int32_t* labels);    /* reset!(o, x); NULL: all */
int32_t stf_get_velocity(stf_ctx* ctx, int32_t n, const int32_t* labels, double* v /*2n*/, uint8_t* has);
/* batchsteps!(qtrees, colist, optimiser)  fit.jl:58-75 in canonical mode: the result equals the
 * sequential loop over `colist` in list order with the draws of stuffing_b200_rng.h keyed
 * (seed, iteration, k).  Steps that share no movable object run concurrently. */
int32_t stf_batchsteps(stf_ctx* ctx, int64_t n, const stf_item* colist, uint64_t seed, uint64_t iteration);
/* one device-resident fit iteration = the body of trainepoch_E!'s rounds (fit.jl:87-96):
 * totalcollisions(qtrees[, labels]; unique=false) followed by batchsteps! on its result, without
 * copying the hit list to the host.  mode 0: spacial over `labels` (NULL: all); mode 1:
 * partialcollisions over `labels`.  *nhits = length(colist). */
int32_t stf_fit_round(stf_ctx* ctx, int32_t mode, int32_t nl, const int32_t* labels,
                      uint64_t seed, uint64_t iteration, int64_t* nhits);
/* the same round, returning the distinct labels of its colist in ascending order instead of the list itself: the
 * `moved := labels of the result` of dynamiccollisions (qtree_functions.jl:354-355) and the `inds` of the element-wise
 * trainers (fit.jl:90: there in order of first appearance) — a dynamic loop then needs ONE call and one synchronisation per
 * step and ships a few hundred labels instead of the hit list.  Two-call pattern on cap / *nmoved. */
int32_t stf_fit_round_moved(stf_ctx* ctx, int32_t mode, int32_t nl, const int32_t* labels, uint64_t seed, uint64_t iteration,
                            int64_t* nhits, int32_t* moved_out, int32_t cap, int32_t* nmoved);
/* the same round followed by getshift(t, level) of EVERY tree (qtrees.jl:135-141): rs/cs (n each) receive the positions
 * after the round.  On the single-scene path the read-back rides on the round's one synchronisation instead of a
 * second call; otherwise it equals stf_fit_round + stf_get_shifts. */
int32_t stf_fit_round_positions(stf_ctx* ctx, int32_t mode, int32_t nl, const int32_t* labels,
                                uint64_t seed, uint64_t iteration, int64_t* nhits,
                                int32_t level, int32_t* rs, int32_t* cs);

/* `n_rounds` fit rounds back to back WITHOUT a host round trip between them — the inner loops of the reference's trainers
 * (trainepoch_E!: fit.jl:91-96; any fixed number of rounds of totalcollisions(qtrees[, labels]) -> batchsteps!) stay on the
 * device: round k uses the draws of iteration0 + k.  The rounds are enqueued blindly; a round that cannot run on the
 * device's current buffers cancels itself and everything after it, the host grows the buffers and continues from that
 * round, so the result always equals n_rounds calls of stf_fit_round.  ONE synchronisation per call on the single-scene
 * latency path.  nhits_per_round[k] = length(colist) of round k; totals (nullable) = sums of items / node tests / hits /
 * moves over the rounds; level > 0: rs/cs (n each) receive getshift(t, level) of every tree after the last round. */
int32_t stf_fit_rounds(stf_ctx* ctx, int32_t mode, int32_t nl, const int32_t* labels, uint64_t seed, uint64_t iteration0,
                       int32_t n_rounds, int64_t* nhits_per_round, stf_counters_t* totals, int32_t level, int32_t* rs, int32_t* cs);

/* trainepoch_E!(qtrees; optimiser)  fit.jl:85-99 with its round loop ON THE DEVICE: round 0 = totalcollisions(qtrees) ->
 * batchsteps!; inds := the labels of its colist in order of first appearance (:90) and the limit length(qtrees) ÷
 * length(inds) are built on the device; rounds ni = 1.. run totalcollisions(qtrees, inds) -> batchsteps! until the round with
 * ni > 8 length(colist) (:96) or ni == limit.  Rounds are enqueued blindly (at most max_rounds per call), ONE
 * synchronisation per call; round k uses the draws of iteration0 + k.  first_round = 0 starts an epoch, > 0 resumes it.
 * state_out = {reason, ni, length(inds), limit}: reason 1 = epoch finished; 0 = max_rounds ran out (call again with
 * first_round = rounds done so far in this epoch); 3 = length(inds) <= 10 = SPACIAL_ENABLE_THRESHOLD (qtree_functions.jl:265,
 * one thread): the dispatcher's native all-pairs path takes over — the caller runs the remaining rounds itself over inds_out.
 * nhits_per_round[0] is the epoch's nc.  Needs a scene on the single-scene latency path. */
int32_t stf_fit_epoch_e(stf_ctx* ctx, uint64_t seed, uint64_t iteration0, int32_t first_round, int32_t max_rounds,
                        int64_t* nhits_per_round, int32_t* rounds_done, int32_t state_out[4], int32_t* inds_out, int32_t inds_cap);

/* ---- ground composition for place! / reposition! (qtree_functions.jl:452-518, fit.jl:420-436).
 * ground = deepcopy(qtrees[mask_label]) with every other tree overlap!-ed except the `excl_labels` (the trees about to be
 * re-placed).  The room search itself (findroom_uniform: a sequential BFS driven by one RNG stream) stays on the host
 * and reads the levels it visits with stf_ground_level (u8 codes, column-major kh x kw; info = kh, kw, rshift, cshift,
 * canvas size of the level). */
int32_t stf_ground_compose(stf_ctx* ctx, int32_t mask_label, int32_t n_excl, const int32_t* excl_labels);
int32_t stf_ground_level(stf_ctx* ctx, int32_t l, uint8_t* out, int32_t info[5]);

/* ---- item-parallel sharding of ONE scene over the GPUs of a box (the reference has no counterpart:
 * its narrow phase is threaded with Threads.@spawn over index chunks, qtree_functions.jl:52-110).
 * Every rank holds an identical replica.  Per fit iteration:
 *   stf_shard_collide : broad phase + narrow phase over the candidate pairs of the sorted records rank, rank+nranks, ...;
 *                       *nlocal = local hit count (the hits stay on the device)
 *   (caller)          : stride = 1 + max over ranks of nlocal (one scalar all-reduce)
 *   stf_shard_pack    : writes [header | nlocal hit records] into send_dev (stride 16-byte records; stream-synchronised)
 *   (caller)          : ONE all-gather of the stride-sized buffers over NVLink (ncclAllGather) into recv_dev
 *   stf_shard_step    : union of the gathered lists -> canonical order -> batchsteps!; every rank computes the same
 *                       moves, so the replicas stay bit-identical without exchanging pyramids.  *nhits = length(colist). */
int32_t stf_shard_collide(stf_ctx* ctx, int32_t mode, int32_t nl, const int32_t* labels, int32_t rank, int32_t nranks, int64_t* nlocal);
int32_t stf_shard_pack(stf_ctx* ctx, void* send_dev, int64_t stride);
int32_t stf_shard_step(stf_ctx* ctx, const void* recv_dev, int32_t nranks, int64_t stride, uint64_t seed, uint64_t iteration, int64_t* nhits);

/* ---- the same sharding for a host that drives SEVERAL GPUs FROM ONE PROCESS (SURVEY §8b threading row: the way a Julia
 * host would; the reference's counterpart is its Threads.@spawn over index chunks, qtree_functions.jl:52-110).
 * stf_group_create makes one context per listed device and enables peer access between them; the caller uploads the SAME
 * scene into every context (stf_group_ctx(g, i) with any stf_upload_*) and applies state changes (stf_set_shifts,
 * stf_set_optimiser, ...) to every context.  stf_group_fit_round = stf_fit_round over the group: GPU i generates and tests
 * the candidate pairs of the sorted records i, i+n, ..., the GPUs exchange their hit lists THEMSELVES (peer stores over
 * NVLink into every peer's gather buffer + a device-side flag barrier, no collective call and no host synchronisation in
 * between), and every GPU runs the same canonical batchsteps on the union, so the replicas stay bit-identical.  Inside the
 * library every GPU has its own (persistent) host thread, so the GPUs are enqueued in parallel; one host synchronisation
 * per GPU and round.  *nhits = length(colist).  Needs a scene that fits the single-scene latency path. */
typedef struct stf_group stf_group;
int32_t stf_group_create(int32_t n, const int32_t* devices, stf_group** out);
void stf_group_destroy(stf_group* g);
int32_t stf_group_size(const stf_group* g);
stf_ctx* stf_group_ctx(stf_group* g, int32_t i);
const char* stf_group_last_error(const stf_group* g);
/* setshift!/shift! (stf_set_shifts with host arrays) and reset!(optimiser) on every replica, the GPUs driven in parallel */
int32_t stf_group_set_shifts(stf_group* g, int32_t n, const int32_t* labels, int32_t level, const int32_t* rs, const int32_t* cs, int32_t mode);
int32_t stf_group_reset_velocity(stf_group* g, int32_t n, const int32_t* labels);
int32_t stf_group_fit_round(stf_group* g, int32_t mode, int32_t nl, const int32_t* labels, uint64_t seed, uint64_t iteration, int64_t* nhits);
/* bytes the last stf_group_fit_round moved between GPUs (hit records and flags pushed to peers) */
int64_t stf_group_nvlink_bytes(const stf_group* g);

int32_t stf_counters(stf_ctx* ctx, stf_counters_t* out);
/* number of kernels this context has launched since stf_create (bench.py's gpu_launches) */
int64_t stf_launch_count(const stf_ctx* ctx);

/* ---- per-stage device timers (CUDA events on the context's stream; for bench.py's roofline) ---- */
#define STF_NSTAGES 9
#define STF_STAGE_SHIFT 0     /* stf_set_shifts: shift update + pyramid rebuild (k_build_objects) */
#define STF_STAGE_LOCATE 1    /* k_locate */
#define STF_STAGE_SORT 2      /* radix sort of the (cell,label) records */
#define STF_STAGE_EMIT 3      /* k_emit_items: candidate generation */
#define STF_STAGE_NARROW 4    /* k_collide_small + k_collide_big: frontier BFS */
#define STF_STAGE_RESULT 5    /* hit compaction, sort!, unique!, export */
#define STF_STAGE_STEP_PREP 6 /* endpoint sort + predecessor links of batchsteps */
#define STF_STAGE_STEP 7      /* k_batchsteps (grad2d + step on the shifts) + k_materialize (one rebuild per moved tree) */
#define STF_STAGE_BUILD 8     /* upload: level-1 pack + full pyramid build from the u8 payload resident in HBM */
int32_t stf_profile_enable(stf_ctx* ctx, int32_t on);
/* accumulated milliseconds and number of timed intervals per stage since the last reset */
int32_t stf_profile_read(stf_ctx* ctx, double* ms, int64_t* calls, int32_t reset);
/* the last upload (profiling on), each build kernel timed alone by its own event pair: ms4 = k_pack_build_small,
 * k_pack_build_mid<128>, k_pack_build_mid<512>, the mask-sized path (k_pack_level1 + wide level builds) */
int32_t stf_profile_build_kernels(stf_ctx* ctx, double* ms4);

#ifdef __cplusplus
}
#endif
#endif
