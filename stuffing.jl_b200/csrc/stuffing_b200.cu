// stuffing_b200.cu — the C ABI (include/stuffing_b200.h) over the sm_100a kernels.
// Host-side orchestration only: buffer management, launches, the few count read-backs the
// reference's own control flow needs (length(itemlist), length(colist)).  No CPU fallback.
#include "../../include/stuffing_b200.h"
#include "stf_common.cuh"
#include "stf_pyramid.cuh"
#include "stf_sort.cuh"
#include "stf_finalize.cuh"
#include "stf_collide.cuh"
#include "stf_step.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes, bool keep = false, cudaStream_t st = 0) {
        if (bytes <= cap) return cudaSuccess;
        size_t ncap = std::max(bytes, cap + cap / 2) + 256;
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, ncap);
        if (e != cudaSuccess) return e;
        if (keep && p && cap) { e = cudaMemcpyAsync(q, p, cap, cudaMemcpyDeviceToDevice, st); if (e == cudaSuccess) e = cudaStreamSynchronize(st); if (e != cudaSuccess) { cudaFree(q); return e; } }
        if (p) cudaFree(p);
        p = q; cap = ncap; return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct stf_ctx {
    int device = 0; cudaStream_t st = nullptr; std::string err;
    int n = 0, L = 0, canvas = 0; uint64_t total_words = 0;
    // environment switches, read ONCE in stf_create (never on the hot path)
    const char* pdl_skip = nullptr; // STF_PDL_SKIP: kernels (by name) launched without the programmatic attribute (debugging)
    bool zero_copy = true;   // STF_NO_ZERO_COPY=1: small host arrays / read-backs go through the copy engine instead of pinned-memory accesses
    bool env_force_big = false; // STF_FORCE_BIG=1: always the multi-launch throughput path
    bool no_cluster = false; // STF_NO_CLUSTER=1: single-CTA sorts instead of the 8-CTA cluster kernels
    bool small_off = false;  // a small scene overflowed the single-CTA sorts once: cluster kernels from then on
    int dbg_step = 0;        // STF_DEBUG_STEP: k_batchsteps experiments
    bool after_kernel = false; // the last operation enqueued on the stream was one of our kernel launches
    bool pdl = true; // programmatic dependent launch between the kernels of a round (STF_NO_PDL=1: plain stream order)
    bool raw_shift_edit = false; // stf_set_level_shift was used: shifts and rasters may disagree, never skip a rebuild
    int scene_bits = 0, label_bits = 1; bool multi_scene = false, sizelevel_ok = true, has_big = false;
    std::vector<ObjMeta> h_om; std::vector<uint32_t> h_off; // h_off[o*L + l-1]: word offsets (static)
    std::vector<int> big_objs;
    DevBuf words, lm, om, payload, poff, woff1, work, lab, lab2, lab3, lab4, bytes_tmp;
    DevBuf rk, rv, rk2, rv2, hist, lkk, lkv;
    DevBuf items, direct, res, overflow, hits, hits2, hk, hv, hk2, hv2, keep, out32, scratch;
    DevBuf built_from, lvlidx, gwords, glm, surv, dbg_t, counts, errflag, vel, has_vel, moved_flag, moved_pos, prev, done, nat_flt, nat_m, dirty;
    unsigned long long* h_counts = nullptr; int* h_err = nullptr; // pinned
    OptState opt = { STF_OPT_MOMENTUM, 0.25, 0.5 };
    int64_t last_result = 0, last_items = 0; bool items_valid = false;
    int64_t big_cap = 0; int big_ctas = 1;
    int ground_of = -1; // object whose copy the ground pyramid currently is (stf_ground_compose)
    int shard_rank = 0, shard_n = 1; // item-parallel sharding of the candidate generation (stf_shard_*, stf_group_*)
    DevBuf gather, gflags, gdone;    // stf_group_*: segments the peers push their hits into, their (count, epoch) flags, push bookkeeping
    bool force_big = false; // the single-CTA latency path overflowed once: stay on the multi-launch path
    int64_t hint_items = 0; // candidate-list size of the previous round (buffer sizing without a host round trip)
    int64_t hint_surv = 0;  // pairs the word-parallel narrow phase handed to the frontier kernel in the previous round
    bool step_global = false;  // STF_STEP_GLOBAL=1: batches of scenes also step through k_batchsteps (global tickets and flags, depth-ordered) instead of k_batchsteps_scene
    bool scene_sorted = false; int n_scenes = 1; // scene ids ascend with the labels: the hits of a scene are contiguous in the canonical order
    DevBuf scene_rng;
    bool sort3 = false;        // STF_SORT_3LAUNCH=1: always the three-launch radix passes and the 16 level sweeps (comparison / tests)
    bool fused_narrow = false; // STF_FUSED_NARROW=1: candidates are tested inside the candidate kernel and never stored (measured slower: DESIGN.md)
    const int32_t* dev_labels = nullptr; const int* dev_nl = nullptr; // stf_fit_epoch_e: the label list of the round lives on the device (the labels of an earlier colist)
    DevBuf inds, epoch_ctl; // that list and the epoch's control block
    bool in_rounds = false; // stf_fit_rounds is enqueueing: k_round_begin resets the counters (not a host memset) and a raised C_STOP cancels a round
    DevBuf lab_mark; uint32_t lab_epoch = 0; bool want_moved = false; // stf_fit_round_moved: epoch-stamped label marks; fold the label list into the round's sync
    DevBuf rounds_rec;      // stf_fit_rounds: per-round records (hits, items, node tests, moves)
    bool keep_items = false; // stf_set_keep_items: the candidate list (the reference's `itemlist`) must be fetchable: never the fused variant
    stf_counters_t ctr = {};
    int64_t launches = 0; // kernels launched by this context (bench.py gpu_launches)
    // per-stage device timers (stf_profile_*): pairs of events recorded on ctx->st
    bool prof = false;
    struct ProfRec { int stage; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_pending; std::vector<cudaEvent_t> prof_pool;
    double prof_ms[STF_NSTAGES] = {}; int64_t prof_calls[STF_NSTAGES] = {};
    int prof_open = -1; cudaEvent_t prof_open_ev = nullptr;
    double build_ms[4] = {}; // last upload, profiling on: k_pack_build_small, k_pack_build_mid<128>, k_pack_build_mid<512>, mask-sized path (each launch timed alone)
    // pinned staging for small host inputs (labels, shifts): the caller's buffer is only read during the call,
    // the H2D copy is a true async DMA, and calls without outputs need no stream synchronisation
    int pos_level = 0; int32_t *pos_rs = nullptr, *pos_cs = nullptr; // stf_fit_round_positions: read-back folded into the round's one sync
    char* h_out = nullptr; size_t h_out_cap = 0; // pinned landing area of small read-backs
    char* stg = nullptr; size_t stg_cap = 0, stg_off = 0; cudaEvent_t stg_ev = nullptr; bool stg_busy = false;
};
// Copies, memsets and event records are not grids: to stay on documented ground, a kernel gets the programmatic attribute
// only when the operation enqueued right before it was one of our kernels.  Every copy/memset/event call in this file
// clears after_kernel.
#define cudaMemsetAsync(...) (ctx->after_kernel = false, cudaMemsetAsync(__VA_ARGS__))
#define cudaMemcpyAsync(...) (ctx->after_kernel = false, cudaMemcpyAsync(__VA_ARGS__))
#define cudaEventRecord(...) (ctx->after_kernel = false, cudaEventRecord(__VA_ARGS__))
// launch with the programmatic-dependent-launch attribute (the kernel must start with pdl_sync()); STF_NO_PDL=1 disables
#define LAUNCH_PDL(ctx, kern, grid, block, smem, ...) do { \
    (ctx)->launches++; \
    if ((ctx)->pdl && (ctx)->after_kernel && !((ctx)->pdl_skip && strstr((ctx)->pdl_skip, #kern))) { \
        cudaLaunchConfig_t cfg_ = {}; cfg_.gridDim = dim3(grid); cfg_.blockDim = dim3(block); cfg_.dynamicSmemBytes = (smem); cfg_.stream = (ctx)->st; \
        cudaLaunchAttribute at_[1]; at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at_[0].val.programmaticStreamSerializationAllowed = 1; \
        cfg_.attrs = at_; cfg_.numAttrs = 1; \
        cudaLaunchKernelEx(&cfg_, kern, __VA_ARGS__); \
    } else kern<<<grid, block, smem, (ctx)->st>>>(__VA_ARGS__); \
    (ctx)->after_kernel = true; } while (0)

// locate: four lanes per label while the labels alone cannot fill the machine, one thread per label beyond that
#define LAUNCH_LOCATE(ctx, nl_, ...) do { \
    if ((int64_t)(nl_) <= 65536) LAUNCH_PDL(ctx, k_locate, nblk(4ll * (nl_), LOC_THREADS), LOC_THREADS, 0, __VA_ARGS__); \
    else LAUNCH_PDL(ctx, k_locate_wide, nblk((nl_), 128), 128, 0, __VA_ARGS__); } while (0)
// several host arrays -> ONE contiguous device range with one copy (dst + sum of the earlier sizes, no padding)
struct HostPart { const void* src; size_t bytes; };
static cudaError_t stage_h2d_parts(stf_ctx* ctx, void* dst, const HostPart* parts, int nparts) {
    size_t bytes = 0;
    for (int i = 0; i < nparts; i++) bytes += parts[i].bytes;
    if (bytes == 0) return cudaSuccess;
    if (ctx->stg_busy && (ctx->stg_off + bytes > ctx->stg_cap || ctx->stg_off == 0)) { cudaEventSynchronize(ctx->stg_ev); ctx->stg_busy = false; ctx->stg_off = 0; }
    if (ctx->stg_off + bytes > ctx->stg_cap) {
        if (ctx->stg_busy) { cudaEventSynchronize(ctx->stg_ev); ctx->stg_busy = false; }
        if (ctx->stg) cudaFreeHost(ctx->stg);
        ctx->stg_cap = std::max(bytes * 2, (size_t)1 << 20); ctx->stg_off = 0;
        cudaError_t e = cudaMallocHost(&ctx->stg, ctx->stg_cap); if (e != cudaSuccess) { ctx->stg = nullptr; ctx->stg_cap = 0; return e; }
    }
    if (!ctx->stg_ev) cudaEventCreateWithFlags(&ctx->stg_ev, cudaEventDisableTiming);
    size_t o = 0;
    for (int i = 0; i < nparts; i++) { memcpy(ctx->stg + ctx->stg_off + o, parts[i].src, parts[i].bytes); o += parts[i].bytes; }
    cudaError_t e = cudaMemcpyAsync(dst, ctx->stg + ctx->stg_off, bytes, cudaMemcpyHostToDevice, ctx->st);
    ctx->stg_off += (bytes + 255) & ~(size_t)255;
    cudaEventRecord(ctx->stg_ev, ctx->st); ctx->stg_busy = true;
    return e;
}
// small inputs: no copy engine at all — the arrays are gathered into the pinned staging area and the consuming kernel reads
// them from there over PCIe (pinned allocations are device-accessible at the same address under unified addressing).
// *dev = that address; the caller launches its kernel and then calls stage_mark() so the area is not reused before it ran.
static cudaError_t stage_zero_copy(stf_ctx* ctx, const HostPart* parts, int nparts, const void** dev) {
    size_t bytes = 0;
    for (int i = 0; i < nparts; i++) bytes += parts[i].bytes;
    if (ctx->stg_busy && (ctx->stg_off + bytes > ctx->stg_cap || ctx->stg_off == 0)) { cudaEventSynchronize(ctx->stg_ev); ctx->stg_busy = false; ctx->stg_off = 0; }
    if (ctx->stg_off + bytes > ctx->stg_cap) {
        if (ctx->stg_busy) { cudaEventSynchronize(ctx->stg_ev); ctx->stg_busy = false; }
        if (ctx->stg) cudaFreeHost(ctx->stg);
        ctx->stg_cap = std::max(bytes * 2, (size_t)1 << 20); ctx->stg_off = 0;
        cudaError_t e = cudaMallocHost(&ctx->stg, ctx->stg_cap); if (e != cudaSuccess) { ctx->stg = nullptr; ctx->stg_cap = 0; return e; }
    }
    if (!ctx->stg_ev) cudaEventCreateWithFlags(&ctx->stg_ev, cudaEventDisableTiming);
    size_t o = 0;
    for (int i = 0; i < nparts; i++) { memcpy(ctx->stg + ctx->stg_off + o, parts[i].src, parts[i].bytes); o += parts[i].bytes; }
    *dev = ctx->stg + ctx->stg_off;
    ctx->stg_off += (bytes + 255) & ~(size_t)255;
    return cudaSuccess;
}
static void stage_mark(stf_ctx* ctx) { cudaEventRecord(ctx->stg_ev, ctx->st); ctx->stg_busy = true; }
static cudaError_t stage_h2d(stf_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return cudaSuccess;
    if (ctx->stg_busy && (ctx->stg_off + bytes > ctx->stg_cap || ctx->stg_off == 0)) { cudaEventSynchronize(ctx->stg_ev); ctx->stg_busy = false; ctx->stg_off = 0; }
    if (ctx->stg_off + bytes > ctx->stg_cap) {
        if (ctx->stg_busy) { cudaEventSynchronize(ctx->stg_ev); ctx->stg_busy = false; }
        if (ctx->stg) cudaFreeHost(ctx->stg);
        ctx->stg_cap = std::max(bytes * 2, (size_t)1 << 20); ctx->stg_off = 0;
        cudaError_t e = cudaMallocHost(&ctx->stg, ctx->stg_cap); if (e != cudaSuccess) { ctx->stg = nullptr; ctx->stg_cap = 0; return e; }
    }
    if (!ctx->stg_ev) cudaEventCreateWithFlags(&ctx->stg_ev, cudaEventDisableTiming);
    memcpy(ctx->stg + ctx->stg_off, src, bytes);
    cudaError_t e = cudaMemcpyAsync(dst, ctx->stg + ctx->stg_off, bytes, cudaMemcpyHostToDevice, ctx->st);
    ctx->stg_off += (bytes + 255) & ~(size_t)255;
    cudaEventRecord(ctx->stg_ev, ctx->st); ctx->stg_busy = true;
    return e;
}
static void stage_reset(stf_ctx* ctx) { // start of an API call: reuse the staging area once the previous copies are done
    if (ctx->stg_busy && cudaEventQuery(ctx->stg_ev) == cudaSuccess) { ctx->stg_busy = false; }
    if (!ctx->stg_busy) ctx->stg_off = 0;
}
static void stage_end(stf_ctx* ctx) {
    if (!ctx->prof || ctx->prof_open < 0) return;
    cudaEvent_t b;
    if (ctx->prof_pool.empty()) cudaEventCreate(&b); else { b = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
    cudaEventRecord(b, ctx->st);
    ctx->prof_pending.push_back({ ctx->prof_open, ctx->prof_open_ev, b });
    ctx->prof_open = -1;
}
static void stage_begin(stf_ctx* ctx, int stage) {
    if (!ctx->prof) return;
    stage_end(ctx);
    cudaEvent_t a;
    if (ctx->prof_pool.empty()) cudaEventCreate(&a); else { a = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
    cudaEventRecord(a, ctx->st);
    ctx->prof_open = stage; ctx->prof_open_ev = a;
}

#define CK(call)                                                                                         \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return STF_E_CUDA; } } while (0)
#define FAIL(code, msg) do { ctx->err = (msg); return (code); } while (0)
#define KCHECK() CK(cudaGetLastError())
static inline int nblk(int64_t n, int t) { return (int)std::max<int64_t>(1, (n + t - 1) / t); }
// entries per warp of the grouped rebuild kernels: one per warp while the GPU has warps to spare, up to 32 beyond that
static inline int rebuild_group(int64_t n) { int64_t g = n / (148 * 64); return g < 4 ? 1 : (int)std::min<int64_t>(32, g); }

enum { C_ITEMS = 0, C_DIRECT = 1, C_OVERFLOW = 2, C_VISITED = 3, C_MOVES = 4, C_KEPT = 5, C_NHITS = 6, C_ZEROED = 7, /* not reset by zero_counts: */ C_NVALID = 7, C_TICKET = 8, C_NEP = 9, C_NSTEPS = 10, C_NREC = 11, C_VARY = 12, C_SURV = 13, C_STATUS = 14 /* path status bits (int) */, C_ERR = 15 /* error bits (int) */, C_CURSOR = 16, C_STOP = 17 /* multi-round calls: != 0 cancels the remaining rounds */, C_N = 18 };
#define ERRP(ctx) reinterpret_cast<int*>((ctx)->counts.as<unsigned long long>() + C_ERR)
#define STATUSP(ctx) reinterpret_cast<int*>((ctx)->counts.as<unsigned long long>() + C_STATUS)

// the counter array -> pinned host memory by posted writes (no copy-engine operation in front of the synchronisation)
__global__ void k_copy_counts(const unsigned long long* __restrict__ counts, unsigned long long* __restrict__ host, int n) {
    pdl_sync();
    if ((int)threadIdx.x < n) host[threadIdx.x] = counts[threadIdx.x];
}
static int32_t sync_counts(stf_ctx* ctx) { // ONE read-back: counters, path status and error bits live in the same array
    if (ctx->zero_copy) { LAUNCH_PDL(ctx, k_copy_counts, 1, 32, 0, (const unsigned long long*)ctx->counts.as<unsigned long long>(), ctx->h_counts, (int)C_N); }
    else CK(cudaMemcpyAsync(ctx->h_counts, ctx->counts.p, C_N * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->h_err[0] = (int)ctx->h_counts[C_ERR]; ctx->h_err[1] = (int)ctx->h_counts[C_STATUS];
    if (ctx->h_err[0]) {
        int e = ctx->h_err[0];
        cudaMemsetAsync(ERRP(ctx), 0, sizeof(unsigned long long), ctx->st);
        if (e & 1) FAIL(STF_E_BADCODE, "node code outside 1..3 in uploaded payload");
        if (e & 2) FAIL(STF_E_RANGE, "spatial cell coordinate outside +-2^21");
        if (e & 4) FAIL(STF_E_CUDA, "internal: frontier scratch too small");
        if (e & 8) FAIL(STF_E_ARG, "device-resident label list: label out of range or mask-sized object");
    }
    return STF_OK;
}
static int32_t upload_labels(stf_ctx* ctx, DevBuf& buf, int32_t nl, const int32_t* labels, const int32_t** dev) {
    *dev = nullptr;
    if (!labels) return STF_OK;
    for (int i = 0; i < nl; i++) if (labels[i] < 1 || labels[i] > ctx->n) FAIL(STF_E_ARG, "label out of range");
    CK(buf.ensure(sizeof(int32_t) * (size_t)std::max(nl, 1)));
    CK(stage_h2d(ctx, buf.p, labels, sizeof(int32_t) * (size_t)nl));
    *dev = buf.as<int32_t>();
    return STF_OK;
}

// ------------------------------------------------------------------------------------------------ lifetime
// Function attributes (> 48 KB of dynamic shared memory, non-portable cluster sizes) are PER DEVICE: they are set for the
// context's device every time a context is created (a host driving several GPUs from one process creates one context per
// device), and a failure fails stf_create instead of surfacing later as a launch error.
static cudaError_t set_func_attrs() {
    cudaError_t e;
#define STF_ATTR(f, a, v) do { e = cudaFuncSetAttribute(f, a, (int)(v)); if (e != cudaSuccess) return e; } while (0)
    const size_t pbm = sizeof(uint32_t) * (PBM_A + PBM_B);
    STF_ATTR(k_build_tail_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, pbm);
    STF_ATTR(k_pack_build_mid<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, pbm);
    STF_ATTR(k_pack_build_mid<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, pbm);
    STF_ATTR(k_sort_records, cudaFuncAttributeMaxDynamicSharedMemorySize, BS_SMEM_BYTES);
    STF_ATTR(k_sort_records_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM_BYTES);
    STF_ATTR(k_finalize_prep_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM_BYTES);
#if CS_CTAS > 8
    STF_ATTR(k_sort_records_cl, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    STF_ATTR(k_finalize_prep_cl, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
#endif
    STF_ATTR(k_finalize_hits, cudaFuncAttributeMaxDynamicSharedMemorySize, BS_SMEM_BYTES);
    STF_ATTR(k_step_links, cudaFuncAttributeMaxDynamicSharedMemorySize, BS_SMEM_BYTES);
#undef STF_ATTR
    return cudaSuccess;
}
extern "C" int32_t stf_create(int32_t device, stf_ctx** out) {
    if (!out) return STF_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return STF_E_CUDA;
    stf_ctx* ctx = new stf_ctx();
    ctx->device = device;
    if (getenv("STF_NO_PDL")) ctx->pdl = false;
    ctx->pdl_skip = getenv("STF_PDL_SKIP");
    ctx->zero_copy = !getenv("STF_NO_ZERO_COPY");
    ctx->env_force_big = getenv("STF_FORCE_BIG") != nullptr;
    ctx->no_cluster = getenv("STF_NO_CLUSTER") != nullptr;
    ctx->fused_narrow = getenv("STF_FUSED_NARROW") != nullptr;
    ctx->sort3 = getenv("STF_SORT_3LAUNCH") != nullptr;
    ctx->step_global = getenv("STF_STEP_GLOBAL") != nullptr;
    { const char* e = getenv("STF_DEBUG_STEP"); ctx->dbg_step = e ? atoi(e) : 0; }
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost(&ctx->h_counts, C_N * sizeof(unsigned long long)) != cudaSuccess || (ctx->h_err = (int*)calloc(2, sizeof(int))) == nullptr ||
        ctx->counts.ensure(C_N * sizeof(unsigned long long)) != cudaSuccess || set_func_attrs() != cudaSuccess) {
        stf_destroy(ctx); return STF_E_CUDA;
    }
    cudaMemsetAsync(ctx->counts.p, 0, C_N * sizeof(unsigned long long), ctx->st);
    *out = ctx;
    return STF_OK;
}
extern "C" void stf_destroy(stf_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    DevBuf* all[] = { &ctx->words, &ctx->lm, &ctx->om, &ctx->payload, &ctx->poff, &ctx->woff1, &ctx->work, &ctx->lab, &ctx->lab2, &ctx->lab3, &ctx->lab4, &ctx->bytes_tmp,
                      &ctx->rk, &ctx->rv, &ctx->rk2, &ctx->rv2, &ctx->hist, &ctx->lkk, &ctx->lkv, &ctx->items, &ctx->direct, &ctx->res, &ctx->overflow,
                      &ctx->hits, &ctx->hits2, &ctx->hk, &ctx->hv, &ctx->hk2, &ctx->hv2, &ctx->keep, &ctx->out32, &ctx->scratch, &ctx->counts, &ctx->errflag,
                      &ctx->vel, &ctx->has_vel, &ctx->moved_flag, &ctx->moved_pos, &ctx->prev, &ctx->done, &ctx->nat_flt, &ctx->nat_m, &ctx->dirty, &ctx->dbg_t, &ctx->surv, &ctx->gwords, &ctx->glm, &ctx->lvlidx, &ctx->built_from, &ctx->gather, &ctx->gflags, &ctx->gdone, &ctx->rounds_rec, &ctx->lab_mark, &ctx->inds, &ctx->epoch_ctl, &ctx->scene_rng };
    for (DevBuf* b : all) b->release();
    if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
    if (ctx->h_err) free(ctx->h_err);
    if (ctx->stg) cudaFreeHost(ctx->stg);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->stg_ev) cudaEventDestroy(ctx->stg_ev);
    if (ctx->st) cudaStreamDestroy(ctx->st);
    delete ctx;
}
extern "C" const char* stf_last_error(const stf_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
extern "C" void* stf_stream(stf_ctx* ctx) { return ctx ? (void*)ctx->st : nullptr; }
extern "C" int32_t stf_num_objects(const stf_ctx* ctx) { return ctx ? ctx->n : 0; }
extern "C" int32_t stf_num_levels(const stf_ctx* ctx) { return ctx ? ctx->L : 0; }

__global__ void k_fill_i32(int* __restrict__ p, int n, int v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

// ------------------------------------------------------------------------------------------------ build helpers
// levels from+1..L of one big object living in (words, lm) at meta index `mi`: wide launches while a level is large, then one CTA
static int32_t build_big_on(stf_ctx* ctx, uint32_t* words, LevelMeta* lm, int mi, int kh1, int kw1, int from) {
    auto words_of = [&](int l) { int kh = ((kh1 - 2) >> (l - 1)) + 2, kw = ((kw1 - 2) >> (l - 1)) + 2; if (l == 1) { kh = kh1; kw = kw1; } return (int64_t)wpc_of(kh) * kw; };
    for (int l = from + 1; l <= ctx->L; l++) {
        int64_t nw = words_of(l);
        if (nw <= PBM_A && (l + 1 > ctx->L || words_of(l + 1) <= PBM_B)) { // level l and everything above by one CTA
            ctx->launches++, k_build_tail_cta<<<1, 512, sizeof(uint32_t) * (PBM_A + PBM_B), ctx->st>>>(words, lm, ctx->L, mi, l - 1, PBM_A);
            break;
        }
        ctx->launches++, k_build_level_wide<<<std::min(nblk(nw, 256), 148 * 16), 256, 0, ctx->st>>>(words, lm, ctx->L, mi, l);
    }
    KCHECK();
    return STF_OK;
}
static int32_t build_big(stf_ctx* ctx, int o, int from) {
    return build_big_on(ctx, ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), o, ctx->h_om[o].kh1, ctx->h_om[o].kw1, from);
}

static int32_t upload_common(stf_ctx* ctx, int32_t n, const uint8_t* const* pics, const int32_t* pitch, const uint8_t* packed, int packed_on_device,
                             const int32_t* kh, const int32_t* kw, const int32_t* rshift, const int32_t* cshift, const uint8_t* dflt, int32_t canvas,
                             const uint8_t* is_mask, const int32_t* scene) {
    if (!ctx) return STF_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (n < 0 || canvas < 2 || (canvas & (canvas - 1)) || canvas > 32768) FAIL(STF_E_ARG, "canvas must be a power of two in [2, 32768]");
    int L = 1; for (int s = canvas; s > 1; s >>= 1) L++;
    ctx->n = n; ctx->L = L; ctx->canvas = canvas; ctx->items_valid = false; ctx->last_result = 0; ctx->raw_shift_edit = false;
    ctx->ground_of = -1; ctx->force_big = false; ctx->small_off = false; ctx->hint_items = 0; ctx->hint_surv = 0; ctx->ctr = stf_counters_t(); // per-scene state of a previous upload
    ctx->label_bits = 1; while ((1ll << ctx->label_bits) <= n) ctx->label_bits++;
    if (n >= (1 << 26)) FAIL(STF_E_ARG, "too many objects");
    ctx->h_om.assign(n, ObjMeta());
    ctx->h_off.assign((size_t)n * L, 0);
    std::vector<LevelMeta> h_lm((size_t)n * L);
    std::vector<uint64_t> poff(n + 1, 0), woff1(n + 1, 0);
    uint64_t words = 0; int max_scene = 0;
    ctx->big_objs.clear(); ctx->sizelevel_ok = true;
    std::vector<int64_t> n2(n, 0);
    std::vector<int32_t> l_small, l_mid, l_mid2, l_big; std::vector<uint64_t> big_woff(1, 0); uint64_t midA[2] = { 0, 0 }, midB[2] = { 0, 0 }; // upload classes (fused pack+build kernels)
    for (int o = 0; o < n; o++) {
        if (kh[o] < 0 || kw[o] < 0 || kh[o] > 32767 || kw[o] > 32767) FAIL(STF_E_ARG, "kernel size must be in [0, 32767]");
        if (dflt[o] != STF_EMPTY && dflt[o] != STF_FULL) FAIL(STF_E_ARG, "default must be EMPTY or FULL");
        ObjMeta& m = ctx->h_om[o];
        m.kh1 = kh[o]; m.kw1 = kw[o]; m.scene = scene ? scene[o] : 0; m.dflt = dflt[o];
        m.is_mask = is_mask ? is_mask[o] : (o == 0);
        if (m.scene < 0) FAIL(STF_E_ARG, "negative scene id");
        max_scene = std::max(max_scene, m.scene);
        int a = kh[o], b = kw[o];
        uint64_t lw[3] = { 0, 0, 0 };
        for (int l = 1; l <= L; l++) {
            LevelMeta& v = h_lm[(size_t)o * L + l - 1];
            if (words > 0xFFFFFFF0ull) FAIL(STF_E_ARG, "pyramids exceed 2^32 words per context");
            v.off = (uint32_t)words; ctx->h_off[(size_t)o * L + l - 1] = v.off;
            v.rs = (l == 1 && rshift) ? rshift[o] : 0; v.cs = (l == 1 && cshift) ? cshift[o] : 0;
            v.dims = (uint32_t)a | ((uint32_t)b << 15) | ((uint32_t)dflt[o] << 30);
            uint64_t nw = (uint64_t)wpc_of(a) * b;
            if (l == 1) { woff1[o + 1] = woff1[o] + nw; poff[o + 1] = poff[o] + (uint64_t)a * b; }
            if (l == 2) { n2[o] = (int64_t)a * b; m.big = nw > 4096; }
            if (l <= 3) lw[l - 1] = nw;
            if (l == L && (a > 2 || b > 2)) ctx->sizelevel_ok = false; // sizelevel == -1  qtrees.jl:182-193
            words += nw;
            a = a / 2 + 1; b = b / 2 + 1; // qtrees.jl:185
        }
        if (m.big) ctx->big_objs.push_back(o);
        if (lw[0] <= RB_HALF) l_small.push_back(o);
        else if (L >= 3 && lw[1] <= PBM_A && lw[2] <= PBM_B && !m.big) { int k = lw[0] <= 4096 ? 0 : 1; (k ? l_mid2 : l_mid).push_back(o); midA[k] = std::max(midA[k], lw[1]); midB[k] = std::max(midB[k], lw[2]); }
        else { l_big.push_back(o); big_woff.push_back(big_woff.back() + lw[0]); }
    }
    ctx->has_big = !ctx->big_objs.empty();
    ctx->total_words = words;
    ctx->multi_scene = max_scene > 0; ctx->n_scenes = max_scene + 1;
    ctx->scene_sorted = true;
    for (int o = 1; o < n && ctx->scene_sorted; o++) if (scene && scene[o] < scene[o - 1]) ctx->scene_sorted = false;
    ctx->scene_bits = 0; while ((1ll << ctx->scene_bits) <= max_scene) ctx->scene_bits++;
    if (!ctx->multi_scene) ctx->scene_bits = 0;
    if (ctx->scene_bits > 64 - 5 - 2 * STF_CELL_W) FAIL(STF_E_ARG, "too many scenes");
    // frontier bound of the large path: a frontier holds distinct nodes that are MIX in BOTH trees, hence at most the
    // level-2 node count of the smaller tree; over all pairs that is the second largest level-2 count
    std::vector<int64_t> s2 = n2; std::sort(s2.begin(), s2.end());
    ctx->big_cap = std::max<int64_t>(1024, n >= 2 ? s2[n - 2] : 1024) + 64;
    ctx->big_ctas = (int)std::max<int64_t>(1, std::min<int64_t>(148, (512ll << 20) / (8 * ctx->big_cap)));

    CK(ctx->words.ensure(sizeof(uint32_t) * (size_t)(words + 16)));
    CK(ctx->lm.ensure(sizeof(LevelMeta) * (size_t)std::max<size_t>(1, (size_t)n * L)));
    CK(ctx->om.ensure(sizeof(ObjMeta) * (size_t)std::max(1, n)));
    CK(ctx->poff.ensure(sizeof(uint64_t) * (size_t)(n + 1)));
    CK(ctx->woff1.ensure(sizeof(uint64_t) * (size_t)(n + 1)));
    CK(ctx->work.ensure(sizeof(int2) * (size_t)std::max(1, n)));
    CK(ctx->vel.ensure(sizeof(double) * 2 * (size_t)std::max(1, n)));
    CK(ctx->has_vel.ensure((size_t)std::max(1, n)));
    CK(ctx->moved_flag.ensure((size_t)std::max(1, n)));
    CK(ctx->moved_pos.ensure(sizeof(int32_t) * (size_t)std::max(1, n)));
    CK(ctx->dirty.ensure(sizeof(int32_t) * (size_t)std::max(1, n)));
    CK(ctx->built_from.ensure(sizeof(int32_t) * (size_t)std::max(1, n)));
    CK(cudaMemsetAsync(ctx->built_from.p, 0, sizeof(int32_t) * (size_t)std::max(1, n), ctx->st)); // the upload builds every tree from level 1
    CK(ctx->lkk.ensure(sizeof(uint64_t) * 4 * (size_t)std::max(1, n)));
    CK(ctx->lkv.ensure(sizeof(uint32_t) * 4 * (size_t)std::max(1, n)));
    CK(cudaMemsetAsync(ctx->has_vel.p, 0, (size_t)std::max(1, n), ctx->st));
    CK(cudaMemsetAsync(ctx->moved_flag.p, 0, (size_t)std::max(1, n), ctx->st));
    if (n == 0) { CK(cudaStreamSynchronize(ctx->st)); return STF_OK; }
    ctx->launches++, k_fill_i32<<<nblk(n, 256), 256, 0, ctx->st>>>(ctx->dirty.as<int>(), n, L); // every raster valid
    CK(cudaMemcpyAsync(ctx->lm.p, h_lm.data(), sizeof(LevelMeta) * (size_t)n * L, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->om.p, ctx->h_om.data(), sizeof(ObjMeta) * (size_t)n, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->poff.p, poff.data(), sizeof(uint64_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, ctx->st));
    ctx->launches++, k_fill_invalid<<<nblk(4ll * n, 256), 256, 0, ctx->st>>>(ctx->lkk.as<uint64_t>(), ctx->lkv.as<uint32_t>(), 4ll * n);

    const uint8_t* d_payload = nullptr;
    std::vector<uint8_t> staging;
    if (packed && packed_on_device) d_payload = packed; // needs >= 32 readable bytes of slack after the payload (zero or valid codes)
    else {
        CK(ctx->payload.ensure((size_t)poff[n] + 64));
        CK(cudaMemsetAsync(ctx->payload.as<uint8_t>() + poff[n], 0, 64, ctx->st)); // pack_word over-reads up to 19 bytes
        const uint8_t* src = packed;
        if (!packed) {
            staging.resize((size_t)poff[n] + 1);
            for (int o = 0; o < n; o++)
                for (int c = 0; c < kw[o]; c++) memcpy(staging.data() + poff[o] + (size_t)c * kh[o], pics[o] + (size_t)c * pitch[o], (size_t)kh[o]);
            src = staging.data();
        }
        if (poff[n]) CK(cudaMemcpyAsync(ctx->payload.p, src, (size_t)poff[n], cudaMemcpyHostToDevice, ctx->st));
        d_payload = ctx->payload.as<uint8_t>();
    }
    // upload lists of the three build classes
    size_t n_mid1 = l_mid.size(); l_mid.insert(l_mid.end(), l_mid2.begin(), l_mid2.end());
    size_t nlist = l_small.size() + l_mid.size() + l_big.size();
    CK(ctx->work.ensure(sizeof(int2) * (size_t)std::max<size_t>(1, std::max<size_t>(n, nlist))));
    CK(ctx->lab4.ensure(sizeof(int32_t) * std::max<size_t>(1, nlist)));
    int32_t* d_list = ctx->lab4.as<int32_t>();
    if (!l_small.empty()) CK(cudaMemcpyAsync(d_list, l_small.data(), sizeof(int32_t) * l_small.size(), cudaMemcpyHostToDevice, ctx->st));
    if (!l_mid.empty()) CK(cudaMemcpyAsync(d_list + l_small.size(), l_mid.data(), sizeof(int32_t) * l_mid.size(), cudaMemcpyHostToDevice, ctx->st));
    if (!l_big.empty()) {
        CK(cudaMemcpyAsync(d_list + l_small.size() + l_mid.size(), l_big.data(), sizeof(int32_t) * l_big.size(), cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(ctx->woff1.p, big_woff.data(), sizeof(uint64_t) * big_woff.size(), cudaMemcpyHostToDevice, ctx->st));
    }
    stage_begin(ctx, STF_STAGE_BUILD);
    cudaEvent_t bev[5] = {};
    auto mark = [&](int i) { if (ctx->prof) { if (!bev[i]) cudaEventCreate(&bev[i]); cudaEventRecord(bev[i], ctx->st); } };
    mark(0);
    // buildqtree!(t) for every object, fused with the level-1 pack: a warp per small object, a CTA per mid-size object;
    // big objects (mask-sized) are packed by a wide launch and built level by level with wide launches
    if (!l_small.empty()) {
        int grp = (int)std::min<int64_t>(32, std::max<int64_t>(4, (int64_t)l_small.size() / (148 * 40 * 2))); // >= 2 groups per resident warp
        int blocks = std::min(nblk(((int64_t)l_small.size() + grp - 1) / grp * 32, 256), 148 * 8);
        ctx->launches++, k_pack_build_small<<<blocks, 256, 0, ctx->st>>>(d_payload, ctx->poff.as<uint64_t>(), d_list, (int)l_small.size(), ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), L, ERRP(ctx), grp);
    }
    mark(1);
    for (int k = 0; k < 2; k++) { // a CTA per mid-size object: 128 threads for a few KB, 512 for mask-sized ones
        size_t cnt = k ? l_mid.size() - n_mid1 : n_mid1;
        if (!cnt) continue;
        int aw = (int)((midA[k] + 31) & ~31ull), bw = (int)((midB[k] + 31) & ~31ull);
        const int32_t* lst = d_list + l_small.size() + (k ? n_mid1 : 0);
        size_t smem = sizeof(uint32_t) * (size_t)(aw + bw);
        if (k == 0) ctx->launches++, k_pack_build_mid<128><<<(int)cnt, 128, smem, ctx->st>>>(d_payload, ctx->poff.as<uint64_t>(), lst, ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), L, aw, ERRP(ctx));
        else ctx->launches++, k_pack_build_mid<512><<<(int)cnt, 512, smem, ctx->st>>>(d_payload, ctx->poff.as<uint64_t>(), lst, ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), L, aw, ERRP(ctx));
        mark(2 + k);
    }
    if (ctx->prof && !bev[2]) { bev[2] = bev[1]; }
    if (ctx->prof && !bev[3]) { bev[3] = bev[2]; }
    if (!l_big.empty()) {
        uint64_t tw = big_woff.back();
        ctx->launches++, k_pack_level1<<<nblk((int64_t)tw, 256), 256, 0, ctx->st>>>(d_payload, ctx->poff.as<uint64_t>(), d_list + l_small.size() + l_mid.size(), ctx->woff1.as<uint64_t>(), (int)l_big.size(),
                                                                      ctx->lm.as<LevelMeta>(), L, ctx->words.as<uint32_t>(), tw, ERRP(ctx));
    }
    KCHECK();
    for (int o : ctx->big_objs) { int32_t r = build_big(ctx, o, 1); if (r) return r; }
    mark(4);
    stage_end(ctx);
    int32_t r = sync_counts(ctx); // staging must outlive the copy; also surfaces bad codes
    if (ctx->prof && !r) {
        for (int i = 0; i < 4; i++) { float t = 0.f; ctx->build_ms[i] = (bev[i] != bev[i + 1] && cudaEventElapsedTime(&t, bev[i], bev[i + 1]) == cudaSuccess) ? t : 0.0; }
    }
    for (int i = 0; i < 5; i++) { bool dup = false; for (int j2 = 0; j2 < i; j2++) dup |= bev[j2] == bev[i]; if (bev[i] && !dup) cudaEventDestroy(bev[i]); }
    return r;
}

extern "C" int32_t stf_upload_objects(stf_ctx* ctx, int32_t n, const uint8_t* const* pics, const int32_t* kh, const int32_t* kw, const int32_t* pitch,
                                      const int32_t* rshift, const int32_t* cshift, const uint8_t* dflt, int32_t canvas, const uint8_t* is_mask, const int32_t* scene) {
    if (!ctx || (n > 0 && (!pics || !kh || !kw || !pitch || !dflt))) return STF_E_ARG;
    return upload_common(ctx, n, pics, pitch, nullptr, 0, kh, kw, rshift, cshift, dflt, canvas, is_mask, scene);
}
extern "C" int32_t stf_upload_packed(stf_ctx* ctx, int32_t n, const uint8_t* payload, int32_t payload_on_device, const int32_t* kh, const int32_t* kw,
                                     const int32_t* rshift, const int32_t* cshift, const uint8_t* dflt, int32_t canvas, const uint8_t* is_mask, const int32_t* scene) {
    if (!ctx || (n > 0 && (!payload || !kh || !kw || !dflt))) return STF_E_ARG;
    return upload_common(ctx, n, nullptr, nullptr, payload, payload_on_device, kh, kw, rshift, cshift, dflt, canvas, is_mask, scene);
}

// ------------------------------------------------------------------------------------------------ tree state
#define NEED_OBJS() do { if (!ctx) return STF_E_ARG; if (ctx->L == 0) FAIL(STF_E_STATE, "nothing uploaded"); CK(cudaSetDevice(ctx->device)); stage_reset(ctx); ctx->after_kernel = false; /* the caller may have put work of its own on the stream */ } while (0)
#define CHECK_LABEL(x) do { if ((x) < 1 || (x) > ctx->n) FAIL(STF_E_ARG, "label out of range"); } while (0)

extern "C" int32_t stf_get_level_info(stf_ctx* ctx, int32_t label, int32_t l, int32_t out[5]) {
    NEED_OBJS(); CHECK_LABEL(label);
    if (l < 1 || l > ctx->L || !out) FAIL(STF_E_ARG, "level out of range");
    LevelMeta m;
    CK(cudaMemcpyAsync(&m, ctx->lm.as<LevelMeta>() + (size_t)(label - 1) * ctx->L + (l - 1), sizeof(m), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    out[0] = lm_kh(m); out[1] = lm_kw(m); out[2] = m.rs; out[3] = m.cs; out[4] = ctx->canvas >> (l - 1);
    return STF_OK;
}
extern "C" int32_t stf_download_level(stf_ctx* ctx, int32_t label, int32_t l, uint8_t* out) {
    NEED_OBJS(); CHECK_LABEL(label);
    if (l < 1 || l > ctx->L) FAIL(STF_E_ARG, "level out of range");
    int o = label - 1;
    int kh = ((ctx->h_om[o].kh1 - 2) >> (l - 1)) + 2, kw = ((ctx->h_om[o].kw1 - 2) >> (l - 1)) + 2;
    if (l == 1) { kh = ctx->h_om[o].kh1; kw = ctx->h_om[o].kw1; }
    size_t nb = (size_t)kh * kw;
    if (!nb) return STF_OK;
    CK(ctx->bytes_tmp.ensure(nb));
    ctx->launches++, k_unpack_level<<<std::min(nblk((int64_t)nb, 256), 148 * 8), 256, 0, ctx->st>>>(ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), ctx->L, o, l, ctx->bytes_tmp.as<uint8_t>());
    KCHECK();
    CK(cudaMemcpyAsync(out, ctx->bytes_tmp.p, nb, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return STF_OK;
}
extern "C" int32_t stf_download_trees(stf_ctx* ctx, int32_t n, const int32_t* labels, uint8_t* out, int64_t cap, int64_t* count) {
    NEED_OBJS();
    if (count) *count = 0;
    if (!labels) n = ctx->n;
    if (n <= 0) return STF_OK;
    const int L = ctx->L;
    std::vector<uint64_t> off((size_t)n * L + 1, 0);
    for (int i = 0; i < n; i++) {
        int lb = labels ? labels[i] : i + 1; CHECK_LABEL(lb);
        int a = ctx->h_om[lb - 1].kh1, b = ctx->h_om[lb - 1].kw1;
        for (int l = 0; l < L; l++) { off[(size_t)i * L + l + 1] = off[(size_t)i * L + l] + (uint64_t)a * b; a = a / 2 + 1; b = b / 2 + 1; }
    }
    const uint64_t total = off.back();
    if (count) *count = (int64_t)total;
    if ((int64_t)total > cap) FAIL(STF_E_CAPACITY, "output buffer too small");
    if (total == 0) return STF_OK;
    if (!out) FAIL(STF_E_ARG, "null output buffer");
    const int32_t* dl; int32_t r = upload_labels(ctx, ctx->lab, n, labels, &dl); if (r) return r;
    CK(ctx->bytes_tmp.ensure((size_t)total)); CK(ctx->woff1.ensure(sizeof(uint64_t) * off.size()));
    CK(cudaMemcpyAsync(ctx->woff1.p, off.data(), sizeof(uint64_t) * off.size(), cudaMemcpyHostToDevice, ctx->st));
    ctx->launches++, k_unpack_trees<<<n * L, 128, 0, ctx->st>>>(ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), L, dl, ctx->woff1.as<uint64_t>(), ctx->bytes_tmp.as<uint8_t>());
    KCHECK();
    CK(cudaMemcpyAsync(out, ctx->bytes_tmp.p, (size_t)total, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return STF_OK;
}
static int32_t upload_query(stf_ctx* ctx, int32_t n, const int32_t* labels, const int32_t* l, const int32_t* a, const int32_t* b) {
    for (int i = 0; i < n; i++) { CHECK_LABEL(labels[i]); if (l[i] < 1 || l[i] > ctx->L) FAIL(STF_E_ARG, "level out of range"); }
    size_t nb = sizeof(int32_t) * (size_t)n;
    CK(ctx->lab.ensure(nb)); CK(ctx->lab2.ensure(nb)); CK(ctx->lab3.ensure(nb)); CK(ctx->lab4.ensure(nb));
    CK(cudaMemcpyAsync(ctx->lab.p, labels, nb, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->lab2.p, l, nb, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->lab3.p, a, nb, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->lab4.p, b, nb, cudaMemcpyHostToDevice, ctx->st));
    return STF_OK;
}
extern "C" int32_t stf_read_nodes(stf_ctx* ctx, int32_t n, const int32_t* labels, const int32_t* l, const int32_t* a, const int32_t* b, uint8_t* out) {
    NEED_OBJS();
    if (n <= 0) return STF_OK;
    int32_t r = upload_query(ctx, n, labels, l, a, b); if (r) return r;
    CK(ctx->bytes_tmp.ensure((size_t)n));
    ctx->launches++, k_read_nodes<<<nblk(n, 256), 256, 0, ctx->st>>>(ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), ctx->L, n, ctx->lab.as<int32_t>(), ctx->lab2.as<int32_t>(),
                                                    ctx->lab3.as<int32_t>(), ctx->lab4.as<int32_t>(), ctx->bytes_tmp.as<uint8_t>());
    KCHECK();
    CK(cudaMemcpyAsync(out, ctx->bytes_tmp.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return STF_OK;
}
extern "C" int32_t stf_grad2d(stf_ctx* ctx, int32_t n, const int32_t* labels, const int32_t* l, const int32_t* a, const int32_t* b, int32_t* out) {
    NEED_OBJS();
    if (n <= 0) return STF_OK;
    int32_t r = upload_query(ctx, n, labels, l, a, b); if (r) return r;
    CK(ctx->out32.ensure(sizeof(int32_t) * 2 * (size_t)n));
    ctx->launches++, k_grad2d<<<nblk(n, 256), 256, 0, ctx->st>>>(ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), ctx->L, n, ctx->lab.as<int32_t>(), ctx->lab2.as<int32_t>(),
                                                ctx->lab3.as<int32_t>(), ctx->lab4.as<int32_t>(), ctx->out32.as<int32_t>());
    KCHECK();
    CK(cudaMemcpyAsync(out, ctx->out32.p, sizeof(int32_t) * 2 * (size_t)n, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return STF_OK;
}
__global__ void k_get_shifts(const LevelMeta* __restrict__ lm, int L, int n, const int32_t* __restrict__ labels, int level, int32_t* __restrict__ out) {
    pdl_sync(); // programmatic dependent launch: scheduled while the previous kernel drains; nothing global before this
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int o = labels ? labels[i] - 1 : i;
    LevelMeta m = lm[(size_t)o * L + (level - 1)];
    out[i] = m.rs; out[n + i] = m.cs;
}
extern "C" int32_t stf_get_shifts(stf_ctx* ctx, int32_t n, const int32_t* labels, int32_t level, int32_t* rs, int32_t* cs) {
    NEED_OBJS();
    if (level < 1 || level > ctx->L || n < 0 || (!labels && n > ctx->n)) FAIL(STF_E_ARG, "bad level or count");
    if (n == 0) return STF_OK;
    const int32_t* dl; int32_t r = upload_labels(ctx, ctx->lab, n, labels, &dl); if (r) return r;
    CK(ctx->out32.ensure(sizeof(int32_t) * 2 * (size_t)n));
    LAUNCH_PDL(ctx, k_get_shifts, nblk(n, 256), 256, 0, ctx->lm.as<LevelMeta>(), ctx->L, n, dl, level, ctx->out32.as<int32_t>());
    KCHECK();
    // one device->host copy into pinned memory, split into the caller's two arrays afterwards
    const size_t nb = sizeof(int32_t) * (size_t)n;
    if (n > 65536) { // large read-backs: straight into the caller's arrays (the extra host copy would cost more than a launch)
        CK(cudaMemcpyAsync(rs, ctx->out32.p, nb, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(cs, ctx->out32.as<int32_t>() + n, nb, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        return STF_OK;
    }
    if (ctx->h_out_cap < 2 * nb) {
        if (ctx->h_out) cudaFreeHost(ctx->h_out);
        ctx->h_out = nullptr; ctx->h_out_cap = 0;
        CK(cudaMallocHost(&ctx->h_out, 4 * nb)); ctx->h_out_cap = 4 * nb;
    }
    CK(cudaMemcpyAsync(ctx->h_out, ctx->out32.p, 2 * nb, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    memcpy(rs, ctx->h_out, nb); memcpy(cs, ctx->h_out + nb, nb);
    return STF_OK;
}
static int32_t rebuild_after(stf_ctx* ctx, int n, const int32_t* labels_host, int from) { // ctx->work filled with (o, from)
    if (from >= ctx->L) return STF_OK;
    int grp = rebuild_group(n);
    int blocks = std::min(nblk(((int64_t)n + grp - 1) / grp * 32, 256), 148 * 8);
    LAUNCH_PDL(ctx, k_build_objects, blocks, 256, 0, ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), ctx->L, ctx->work.as<int2>(), n, ctx->om.as<ObjMeta>(), ctx->has_big ? 1 : 0, grp,
                                                                 nullptr, nullptr, nullptr, 0, 0, ctx->built_from.as<int>(), 0, 0, nullptr);
    KCHECK();
    if (ctx->has_big)
        for (int i = 0; i < n; i++) { int o = labels_host ? labels_host[i] - 1 : i; if (ctx->h_om[o].big) { int32_t r = build_big(ctx, o, from); if (r) return r; } }
    return STF_OK;
}
extern "C" int32_t stf_set_shifts(stf_ctx* ctx, int32_t n, const int32_t* labels, int32_t level, const int32_t* rs, const int32_t* cs, int32_t mode) {
    NEED_OBJS();
    if (level < 1 || level > ctx->L || n < 0 || (!labels && n > ctx->n)) FAIL(STF_E_ARG, "bad level or count");
    if ((mode & STF_SHIFT_DEVICE_LABELS) && !(mode & STF_SHIFT_DEVICE_ARRAYS)) FAIL(STF_E_ARG, "a device label list needs device rs/cs arrays");
    if (n == 0) return STF_OK;
    stage_begin(ctx, STF_STAGE_SHIFT);
    const int32_t* dl = nullptr; int32_t r;
    const int32_t *drs = rs, *dcs = cs;
    bool zero_copy = false;
    if (!(mode & STF_SHIFT_DEVICE_ARRAYS)) { // host arrays: labels, rs and cs travel in ONE copy
        if (labels) for (int i = 0; i < n; i++) if (labels[i] < 1 || labels[i] > ctx->n) FAIL(STF_E_ARG, "label out of range");
        const size_t nb = sizeof(int32_t) * (size_t)n;
        HostPart parts[3] = { { labels, labels ? nb : 0 }, { rs, nb }, { cs, nb } };
        const int32_t* base;
        zero_copy = n <= 65536 && ctx->zero_copy; // up to 768 KB: read by the kernel straight from pinned memory
        if (zero_copy) { const void* zp = nullptr; CK(stage_zero_copy(ctx, parts, 3, &zp)); base = static_cast<const int32_t*>(zp); }
        else { CK(ctx->lab.ensure(3 * nb)); CK(stage_h2d_parts(ctx, ctx->lab.p, parts, 3)); base = ctx->lab.as<int32_t>(); }
        if (labels) { dl = base; base += n; }
        drs = base; dcs = base + n;
    } else if (mode & STF_SHIFT_DEVICE_LABELS) dl = labels; // resident list: validated by the kernel
    else { r = upload_labels(ctx, ctx->lab, n, labels, &dl); if (r) return r; }
    const bool dev_labels = (mode & STF_SHIFT_DEVICE_LABELS) && (mode & STF_SHIFT_DEVICE_ARRAYS) && labels;
    { // one launch: shift + rebuild of every listed tree (mask-sized ones: shift here, rebuild by the wide launches below)
        int grp = rebuild_group(n);
        int blocks = std::min(nblk(((int64_t)n + grp - 1) / grp * 32, 256), 148 * 8);
        LAUNCH_PDL(ctx, k_build_objects, blocks, 256, 0, ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), ctx->L, nullptr, n, ctx->om.as<ObjMeta>(), ctx->has_big ? 1 : 0, grp,
                                                                     dl, drs, dcs, level, mode & 1, ctx->built_from.as<int>(), ctx->raw_shift_edit ? 0 : 1, dev_labels ? ctx->n : 0, ERRP(ctx));
        KCHECK();
        if (zero_copy) stage_mark(ctx); // the staging area is free again once this kernel has run
        if (ctx->has_big && level < ctx->L && !dev_labels)
            for (int i = 0; i < n; i++) { int o = labels ? labels[i] - 1 : i; if (ctx->h_om[o].big) { r = build_big(ctx, o, level); if (r) return r; } }
    }
    stage_end(ctx);
    return STF_OK; // stream-ordered: no output, inputs were staged
}
extern "C" int32_t stf_set_level_shift(stf_ctx* ctx, int32_t label, int32_t l, int32_t rs, int32_t cs) {
    NEED_OBJS(); CHECK_LABEL(label);
    if (l < 1 || l > ctx->L) FAIL(STF_E_ARG, "level out of range");
    int32_t v[2] = { rs, cs };
    ctx->raw_shift_edit = true;
    CK(stage_h2d(ctx, reinterpret_cast<char*>(ctx->lm.as<LevelMeta>() + (size_t)(label - 1) * ctx->L + (l - 1)) + 4, v, 8));
    return STF_OK;
}
extern "C" int32_t stf_build(stf_ctx* ctx, int32_t n, const int32_t* labels, int32_t layer) {
    NEED_OBJS();
    if (layer < 2) FAIL(STF_E_ARG, "layer must be >= 2");
    if (!labels) n = ctx->n;
    if (n == 0 || layer > ctx->L) return STF_OK;
    const int32_t* dl; int32_t r = upload_labels(ctx, ctx->lab, n, labels, &dl); if (r) return r;
    CK(ctx->work.ensure(sizeof(int2) * (size_t)n));
    ctx->launches++, k_fill_worklist<<<nblk(n, 256), 256, 0, ctx->st>>>(n, dl, layer - 1, ctx->work.as<int2>());
    r = rebuild_after(ctx, n, labels, layer - 1); if (r) return r;
    return STF_OK;
}

// ------------------------------------------------------------------------------------------------ narrow phase
// survivor / overflow lists of the narrow phase (SurvRec, 32 B): sized for `want` pairs
static int32_t ensure_surv(stf_ctx* ctx, int64_t want) {
    CK(ctx->surv.ensure(sizeof(SurvRec) * (size_t)std::max<int64_t>(1024, want)));
    CK(ctx->overflow.ensure(sizeof(SurvRec) * (size_t)std::max<int64_t>(1024, want)));
    return STF_OK;
}
static void fill_narrow_args(stf_ctx* ctx, NarrowArgs& A, bool want_res, bool want_hits, bool device_counts) {
    A.words = ctx->words.as<uint32_t>(); A.lm = ctx->lm.as<LevelMeta>(); A.L = ctx->L;
    A.items = ctx->items.as<Item16>(); A.n_items_dev = nullptr; A.n_items_host = 0;
    A.res = want_res ? ctx->res.as<int4>() : nullptr; A.counts = ctx->counts.as<unsigned long long>();
    A.hits = want_hits ? ctx->hits.as<Item16>() : nullptr; A.hits_cap = (int64_t)(ctx->hits.cap / sizeof(Item16)); A.nhits = ctx->counts.as<unsigned long long>() + C_NHITS;
    A.overflow = ctx->overflow.as<SurvRec>(); A.overflow_cap = (int64_t)(ctx->overflow.cap / sizeof(SurvRec));
    A.surv = ctx->surv.as<SurvRec>(); A.surv_cap = (int64_t)(ctx->surv.cap / sizeof(SurvRec));
    A.nsurv = ctx->counts.as<unsigned long long>() + C_SURV; A.cursor = ctx->counts.as<unsigned long long>() + C_CURSOR;
    A.status = device_counts ? STATUSP(ctx) : nullptr;
}
// the frontier kernels over the survivor list (filled by k_narrow_items or by the fused candidate kernel): `want` = expected
// number of survivors (grid sizing only; the kernels read the real count from the device)
static int32_t launch_frontier(stf_ctx* ctx, const NarrowArgs& A, int64_t want) {
    // few pairs (a single scene): a whole warp per pair shortens the dependent chain of the deep ones; many pairs (a batch):
    // four lanes per pair for throughput
    if (want <= 32768) {
        int blocks = (int)std::min<int64_t>(std::max<int64_t>(1, (want + 3) / 4), 148 * 16);
        LAUNCH_PDL(ctx, k_collide_small<32>, blocks, NP_THREADS, 0, A);
    } else {
        int per_block = NP_THREADS / 4;
        int blocks = (int)std::min<int64_t>(std::max<int64_t>(1, (want + per_block - 1) / per_block), 148 * 16); // the warps share the list dynamically
        LAUNCH_PDL(ctx, k_collide_small<4>, blocks, NP_THREADS, 0, A);
    }
    KCHECK();
    CK(ctx->scratch.ensure(sizeof(uint32_t) * 2 * (size_t)ctx->big_cap * ctx->big_ctas));
    LAUNCH_PDL(ctx, k_collide_big, ctx->big_ctas, NPB_THREADS, 0, A, ctx->scratch.as<uint32_t>(), ctx->big_cap, ERRP(ctx));
    KCHECK();
    return STF_OK;
}
// runs the narrow phase over the explicit list ctx->items.  n_host >= 0: host-known item count; n_host < 0: the count is on
// the device (counts[C_ITEMS], clamped to `cap`).  want_res: per-item results into ctx->res; want_hits: hits (l >= 0)
// appended to ctx->hits through counts[C_NHITS].
static int32_t narrow_phase(stf_ctx* ctx, int64_t n_host, int64_t cap, bool want_res, bool want_hits) {
    int64_t nmax = n_host >= 0 ? n_host : cap;
    if (want_res) CK(ctx->res.ensure(sizeof(int4) * (size_t)std::max<int64_t>(1, nmax)));
    if (nmax == 0) return STF_OK;
    if (nmax >= (1ll << 32)) FAIL(STF_E_ARG, "more than 2^32 candidate pairs in one call");
    int32_t r = ensure_surv(ctx, nmax); if (r) return r; // every item may survive: never too small
    NarrowArgs A; fill_narrow_args(ctx, A, want_res, want_hits, n_host < 0);
    A.n_items_dev = n_host >= 0 ? nullptr : ctx->counts.as<unsigned long long>() + C_ITEMS; A.n_items_host = nmax;
    if (n_host >= 0) CK(cudaMemsetAsync(A.nsurv, 0, sizeof(unsigned long long), ctx->st)); // (the latency path zeroes the whole counter array once)
    int64_t want = n_host >= 0 ? n_host : std::max<int64_t>(ctx->hint_items, 4096);
    LAUNCH_PDL(ctx, k_narrow_items, (int)std::min<int64_t>(std::max<int64_t>(1, (want + NW_THREADS - 1) / NW_THREADS), 148 * 8), NW_THREADS, 0, A);
    return launch_frontier(ctx, A, std::max<int64_t>(want / 16, 256));
}
static int32_t import_items(stf_ctx* ctx, int64_t n, const stf_item* items, DevBuf& dst) {
    for (int64_t i = 0; i < n; i++) {
        CHECK_LABEL(items[i].i1); CHECK_LABEL(items[i].i2);
        if (items[i].l < 1 || items[i].l > ctx->L) FAIL(STF_E_ARG, "item level out of range");
    }
    CK(ctx->out32.ensure(sizeof(int32_t) * 5 * (size_t)std::max<int64_t>(1, n)));
    CK(dst.ensure(sizeof(Item16) * (size_t)std::max<int64_t>(1, n)));
    if (n == 0) return STF_OK;
    CK(cudaMemcpyAsync(ctx->out32.p, items, sizeof(stf_item) * (size_t)n, cudaMemcpyHostToDevice, ctx->st));
    ctx->launches++, k_import_items<<<nblk(n, 256), 256, 0, ctx->st>>>(ctx->out32.as<int32_t>(), n, dst.as<Item16>());
    KCHECK();
    return STF_OK;
}
static int32_t zero_counts(stf_ctx* ctx) { CK(cudaMemsetAsync(ctx->counts.p, 0, C_ZEROED * sizeof(unsigned long long), ctx->st)); return STF_OK; }

extern "C" int32_t stf_collision_bfs(stf_ctx* ctx, int64_t n, const stf_item* items, stf_item* out) {
    NEED_OBJS();
    if (n <= 0) return STF_OK;
    for (int64_t i = 0; i < n; i++) if (items[i].l < 2) FAIL(STF_E_ARG, "_collision_randbfs needs a start level >= 2");
    int32_t r = import_items(ctx, n, items, ctx->items); if (r) return r;
    ctx->items_valid = false;
    r = zero_counts(ctx); if (r) return r;
    r = narrow_phase(ctx, n, n, true, false); if (r) return r;
    ctx->launches++, k_export_results<<<nblk(n, 256), 256, 0, ctx->st>>>(ctx->items.as<Item16>(), ctx->res.as<int4>(), n, ctx->out32.as<int32_t>());
    KCHECK();
    CK(cudaMemcpyAsync(out, ctx->out32.p, sizeof(stf_item) * (size_t)n, cudaMemcpyDeviceToHost, ctx->st));
    r = sync_counts(ctx); if (r) return r;
    ctx->ctr.items = n; ctx->ctr.direct = 0; ctx->ctr.hits = 0;
    ctx->ctr.overflow = (int64_t)ctx->h_counts[C_OVERFLOW]; ctx->ctr.visited = (int64_t)ctx->h_counts[C_VISITED];
    return STF_OK;
}
extern "C" int32_t stf_collision_dfs(stf_ctx* ctx, int64_t n, const stf_item* items, stf_item* out) {
    NEED_OBJS();
    if (n <= 0) return STF_OK;
    int32_t r = import_items(ctx, n, items, ctx->items); if (r) return r;
    ctx->items_valid = false;
    CK(ctx->res.ensure(sizeof(int4) * (size_t)n));
    ctx->launches++, k_collision_dfs<<<nblk(n, 128), 128, 0, ctx->st>>>(ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), ctx->L, ctx->items.as<Item16>(), n, ctx->res.as<int4>());
    KCHECK();
    ctx->launches++, k_export_results<<<nblk(n, 256), 256, 0, ctx->st>>>(ctx->items.as<Item16>(), ctx->res.as<int4>(), n, ctx->out32.as<int32_t>());
    KCHECK();
    CK(cudaMemcpyAsync(out, ctx->out32.p, sizeof(stf_item) * (size_t)n, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return STF_OK;
}
extern "C" int32_t stf_collision(stf_ctx* ctx, int32_t i1, int32_t i2, int32_t out[3]) {
    NEED_OBJS(); CHECK_LABEL(i1); CHECK_LABEL(i2);
    // inkernelbounds(Q, L, 1, 1) for both, else (-L, 1, 1)   qtree_functions.jl:46-50
    int32_t a[5], b[5];
    int32_t r = stf_get_level_info(ctx, i1, ctx->L, a); if (r) return r;
    r = stf_get_level_info(ctx, i2, ctx->L, b); if (r) return r;
    auto inb = [](const int32_t* m) { return 1 > m[2] && 1 > m[3] && 1 <= m[0] + m[2] && 1 <= m[1] + m[3]; };
    if (!(inb(a) && inb(b))) { out[0] = -ctx->L; out[1] = 1; out[2] = 1; return STF_OK; }
    stf_item it = { i1, i2, ctx->L, 1, 1 }, res;
    r = stf_collision_bfs(ctx, 1, &it, &res); if (r) return r;
    out[0] = res.l; out[1] = res.a; out[2] = res.b;
    return STF_OK;
}

// ------------------------------------------------------------------------------------------------ result pipeline
// ctx->hits[0..nh) (host-known count) → sorted by ((i1,i2),(l,a,b)) in ctx->hits2, optional unique, exported to ctx->out32
static int32_t finalize_hits(stf_ctx* ctx, int64_t nh, int unique, bool export_host_format, int64_t* nout) {
    *nout = 0;
    CK(ctx->hits2.ensure(sizeof(Item16) * (size_t)std::max<int64_t>(1, nh)));
    if (nh == 0) return STF_OK;
    stage_begin(ctx, STF_STAGE_RESULT);
    if (nh <= BS_CAP && 2 * ctx->label_bits + FIN_IDX_BITS <= 64) { // one CTA: sort, tie order, unique, export
        if (export_host_format) CK(ctx->out32.ensure(sizeof(int32_t) * 5 * (size_t)nh));
        unsigned long long v = (unsigned long long)nh;
        CK(stage_h2d(ctx, ctx->counts.as<unsigned long long>() + C_NHITS, &v, sizeof(v)));
        FinalizeArgs F = {};
        F.hits = ctx->hits.as<Item16>(); F.nh_dev = ctx->counts.as<unsigned long long>() + C_NHITS; F.hits_cap = nh;
        F.n_items_dev = nullptr; F.item_cap = 0; F.sorted = ctx->hits2.as<Item16>(); F.label_bits = ctx->label_bits; F.unique = unique;
        F.out5 = export_host_format ? ctx->out32.as<int32_t>() : nullptr; F.kept = ctx->counts.as<unsigned long long>() + C_KEPT;
        F.do_prep = 0; F.status = STATUSP(ctx);
        LAUNCH_PDL(ctx, k_finalize_hits, 1, BS_THREADS, BS_SMEM_BYTES, F);
        KCHECK();
        stage_end(ctx);
        if (!export_host_format) { *nout = nh; return STF_OK; }
        int32_t r = sync_counts(ctx); if (r) return r;
        *nout = (int64_t)ctx->h_counts[C_KEPT];
        return STF_OK;
    }
    size_t kb = sizeof(uint64_t) * (size_t)nh, vb = sizeof(uint32_t) * (size_t)nh;
    CK(ctx->hk.ensure(kb)); CK(ctx->hk2.ensure(kb)); CK(ctx->hv.ensure(vb)); CK(ctx->hv2.ensure(vb)); CK(ctx->keep.ensure(vb + 16));
    CK(ctx->hist.ensure(sizeof(uint32_t) * radix_sort_hist_words(nh)));
    LAUNCH_PDL(ctx, k_hit_keys, nblk(nh, 256), 256, 0, ctx->hits.as<Item16>(), nh, ctx->hk.as<uint64_t>(), ctx->hv.as<uint32_t>());
    uint64_t* ks; uint32_t* vs;
    { // (i1, i2) = bits [32, 32+lb) and [0, lb): one LSD sort over exactly those bits
        int f = radix_sort_pairs(ctx->st, &ctx->launches, ctx->hk.as<uint64_t>(), ctx->hv.as<uint32_t>(), ctx->hk2.as<uint64_t>(), ctx->hv2.as<uint32_t>(), nh,
                                 bit_range(0, ctx->label_bits) | bit_range(32, 32 + ctx->label_bits), ctx->hist.as<uint32_t>(), ctx->pdl, &ctx->after_kernel, ctx->sort3);
        ks = f ? ctx->hk2.as<uint64_t>() : ctx->hk.as<uint64_t>(); vs = f ? ctx->hv2.as<uint32_t>() : ctx->hv.as<uint32_t>();
    }
    LAUNCH_PDL(ctx, k_gather_hits, nblk(nh, 256), 256, 0, ctx->hits.as<Item16>(), vs, nh, ctx->hits2.as<Item16>());
    LAUNCH_PDL(ctx, k_fix_ties, nblk(nh, 256), 256, 0, ctx->hits2.as<Item16>(), ks, nh, unique, ctx->keep.as<uint32_t>());
    KCHECK();
    if (!export_host_format && !unique) { *nout = nh; stage_end(ctx); return STF_OK; }
    // positions = exclusive scan of keep (reuse the other value buffer)
    uint32_t* pos = (vs == ctx->hv.as<uint32_t>()) ? ctx->hv2.as<uint32_t>() : ctx->hv.as<uint32_t>();
    CK(cudaMemcpyAsync(pos, ctx->keep.p, vb, cudaMemcpyDeviceToDevice, ctx->st));
    LAUNCH_PDL(ctx, k_scan_u32, 1, 1024, 0, pos, nh, reinterpret_cast<uint32_t*>(ctx->counts.as<unsigned long long>() + C_KEPT));
    CK(ctx->out32.ensure(sizeof(int32_t) * 5 * (size_t)nh));
    LAUNCH_PDL(ctx, k_export_hits, nblk(nh, 256), 256, 0, ctx->hits2.as<Item16>(), ctx->keep.as<uint32_t>(), pos, nh, ctx->out32.as<int32_t>());
    KCHECK();
    stage_end(ctx);
    int32_t r = sync_counts(ctx); if (r) return r;
    *nout = (int64_t)(uint32_t)ctx->h_counts[C_KEPT];
    return STF_OK;
}
static int32_t deliver(stf_ctx* ctx, int64_t n, stf_item* out, int64_t cap, int64_t* count) {
    ctx->last_result = n;
    if (count) *count = n;
    if (n > cap) FAIL(STF_E_CAPACITY, "output buffer too small");
    if (n > 0) {
        if (!out) FAIL(STF_E_ARG, "null output buffer");
        CK(cudaMemcpyAsync(out, ctx->out32.p, sizeof(stf_item) * (size_t)n, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
    }
    return STF_OK;
}
extern "C" int32_t stf_fetch_result(stf_ctx* ctx, stf_item* out, int64_t cap, int64_t* count) {
    NEED_OBJS();
    return deliver(ctx, ctx->last_result, out, cap, count);
}
extern "C" int32_t stf_set_keep_items(stf_ctx* ctx, int32_t on) {
    if (!ctx) return STF_E_ARG;
    ctx->keep_items = on != 0; ctx->items_valid = false; // (only matters with STF_FUSED_NARROW=1: the default path always materialises the list)
    return STF_OK;
}
extern "C" int32_t stf_fetch_items(stf_ctx* ctx, stf_item* out, int64_t cap, int64_t* count) {
    NEED_OBJS();
    if (!ctx->items_valid) FAIL(STF_E_STATE, "no item list available: call stf_set_keep_items(ctx, 1) before the many-body call");
    int64_t n = ctx->last_items;
    if (count) *count = n;
    if (n > cap) FAIL(STF_E_CAPACITY, "output buffer too small");
    if (n == 0) return STF_OK;
    CK(ctx->bytes_tmp.ensure(sizeof(int32_t) * 5 * (size_t)n));
    ctx->launches++, k_export_items<<<nblk(n, 256), 256, 0, ctx->st>>>(ctx->items.as<Item16>(), n, ctx->bytes_tmp.as<int32_t>());
    KCHECK();
    CK(cudaMemcpyAsync(out, ctx->bytes_tmp.p, sizeof(stf_item) * (size_t)n, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return STF_OK;
}

// items in ctx->items (n known on host), `nd` direct hits already in ctx->hits[0..nd): BFS, hits appended; *nh = total
static int32_t collide_and_collect(stf_ctx* ctx, int64_t n, int64_t nd, int64_t* nh) {
    CK(ctx->hits.ensure(sizeof(Item16) * (size_t)std::max<int64_t>(1, n + nd), true, ctx->st));
    unsigned long long base = (unsigned long long)nd;
    CK(stage_h2d(ctx, ctx->counts.as<unsigned long long>() + C_NHITS, &base, sizeof(base)));
    stage_begin(ctx, STF_STAGE_NARROW);
    int32_t r = narrow_phase(ctx, n, n, false, true); if (r) return r;
    stage_end(ctx);
    r = sync_counts(ctx); if (r) return r;
    *nh = (int64_t)ctx->h_counts[C_NHITS];
    ctx->ctr.items = n; ctx->ctr.direct = nd; ctx->ctr.hits = *nh;
    ctx->ctr.overflow = (int64_t)ctx->h_counts[C_OVERFLOW]; ctx->ctr.visited = (int64_t)ctx->h_counts[C_VISITED];
    ctx->last_items = n; ctx->items_valid = true; ctx->hint_items = n;
    return STF_OK;
}

// filttrain!'s pool test on the device (fit.jl:235-266): the pool is `pairs` (n_pairs x 2 labels, host) or —
// pairs == NULL — ALL pairs i < j of the objects, generated on the device in the reference's order (fit.jl:279: at 10 k
// objects that is 5e7 pairs, never enumerated or shipped by the host).  Every pair is tested from the root (L,1,1); the
// pairs with result level >= nearlevel come back as (i1,i2) => (l,a,b) in pool order: level > 0 = a collision to step,
// level in [nearlevel, 0) = a near miss for the next pool.
extern "C" int32_t stf_filter_pairs(stf_ctx* ctx, int64_t n_pairs, const int32_t* pairs, int32_t nearlevel, stf_item* out, int64_t cap, int64_t* count) {
    NEED_OBJS();
    if (count) *count = 0;
    ctx->last_result = 0;
    if (ctx->multi_scene) FAIL(STF_E_STATE, "pair pools are defined for one scene");
    const int64_t n = pairs ? n_pairs : (int64_t)ctx->n * (ctx->n - 1) / 2;
    if (n <= 0) return STF_OK;
    if (ctx->L < 2) FAIL(STF_E_STATE, "_collision_randbfs needs a start level >= 2");
    CK(ctx->items.ensure(sizeof(Item16) * (size_t)n));
    ctx->items_valid = false;
    if (pairs) {
        for (int64_t i = 0; i < 2 * n; i++) CHECK_LABEL(pairs[i]);
        CK(ctx->out32.ensure(sizeof(int32_t) * 2 * (size_t)n));
        CK(cudaMemcpyAsync(ctx->out32.p, pairs, sizeof(int32_t) * 2 * (size_t)n, cudaMemcpyHostToDevice, ctx->st));
        ctx->launches++, k_pairs_items<<<nblk(n, 256), 256, 0, ctx->st>>>(ctx->out32.as<int32_t>(), n, ctx->L, ctx->items.as<Item16>());
    } else {
        ctx->launches++, k_allpairs_items<<<nblk((int64_t)ctx->n * ctx->n, 256), 256, 0, ctx->st>>>(ctx->n, ctx->L, ctx->items.as<Item16>());
    }
    KCHECK();
    int32_t r = zero_counts(ctx); if (r) return r;
    r = narrow_phase(ctx, n, n, true, false); if (r) return r;
    const int nb = nblk(n, FP_BLOCK);
    CK(ctx->keep.ensure(sizeof(uint32_t) * (size_t)(nb + 16)));
    ctx->launches++, k_near_count<<<nb, FP_BLOCK, 0, ctx->st>>>(ctx->res.as<int4>(), n, nearlevel, ctx->keep.as<uint32_t>());
    ctx->after_kernel = false;
    LAUNCH_PDL(ctx, k_scan_u32, 1, 1024, 0, ctx->keep.as<uint32_t>(), (int64_t)nb, reinterpret_cast<uint32_t*>(ctx->counts.as<unsigned long long>() + C_KEPT));
    KCHECK();
    r = sync_counts(ctx); if (r) return r;
    const int64_t kept = (int64_t)(uint32_t)ctx->h_counts[C_KEPT];
    ctx->ctr.items = n; ctx->ctr.direct = 0; ctx->ctr.hits = 0; ctx->ctr.overflow = (int64_t)ctx->h_counts[C_OVERFLOW]; ctx->ctr.visited = (int64_t)ctx->h_counts[C_VISITED];
    CK(ctx->out32.ensure(sizeof(int32_t) * 5 * (size_t)std::max<int64_t>(1, kept)));
    if (kept > 0) {
        ctx->launches++, k_near_write<<<nb, FP_BLOCK, 0, ctx->st>>>(ctx->items.as<Item16>(), ctx->res.as<int4>(), n, nearlevel, ctx->keep.as<uint32_t>(), ctx->out32.as<int32_t>());
        KCHECK();
    }
    return deliver(ctx, kept, out, cap, count);
}

extern "C" int32_t stf_collide_items(stf_ctx* ctx, int64_t n, const stf_item* items, stf_item* out, int64_t cap, int64_t* count) {
    NEED_OBJS();
    if (count) *count = 0;
    if (n <= 0) { ctx->last_result = 0; return STF_OK; }
    for (int64_t i = 0; i < n; i++) if (items[i].l < 2) FAIL(STF_E_ARG, "_collision_randbfs needs a start level >= 2");
    int32_t r = import_items(ctx, n, items, ctx->items); if (r) return r;
    r = zero_counts(ctx); if (r) return r;
    int64_t nh = 0, nout = 0;
    r = collide_and_collect(ctx, n, 0, &nh); if (r) return r;
    r = finalize_hits(ctx, nh, 0, true, &nout); if (r) return r;
    return deliver(ctx, nout, out, cap, count);
}

extern "C" int32_t stf_totalcollisions_native(stf_ctx* ctx, int32_t nl, const int32_t* labels, stf_item* out, int64_t cap, int64_t* count) {
    NEED_OBJS();
    if (count) *count = 0;
    ctx->last_result = 0;
    if (ctx->multi_scene) FAIL(STF_E_STATE, "totalcollisions_native is defined for one scene");
    if (!labels) nl = ctx->n;
    if (ctx->n <= 1 || nl <= 1) return STF_OK; // length(qtrees) > 1 || return colist  :127
    const int32_t* dl; int32_t r = upload_labels(ctx, ctx->lab, nl, labels, &dl); if (r) return r;
    r = zero_counts(ctx); if (r) return r;
    int64_t npairs = (int64_t)nl * (nl - 1) / 2;
    CK(ctx->nat_flt.ensure(sizeof(int32_t) * (size_t)nl)); CK(ctx->nat_m.ensure(sizeof(int)));
    CK(ctx->items.ensure(sizeof(Item16) * (size_t)npairs));
    ctx->launches++, k_native_filter<<<1, 1024, 0, ctx->st>>>(ctx->lm.as<LevelMeta>(), ctx->L, nl, dl, ctx->nat_flt.as<int32_t>(), ctx->nat_m.as<int>());
    ctx->launches++, k_native_pairs<<<nblk((int64_t)nl * nl, 256), 256, 0, ctx->st>>>(ctx->nat_flt.as<int32_t>(), ctx->nat_m.as<int>(), nl, ctx->L, ctx->items.as<Item16>(), npairs,
                                                                    ctx->counts.as<unsigned long long>());
    KCHECK();
    r = sync_counts(ctx); if (r) return r;
    int64_t n = (int64_t)ctx->h_counts[C_ITEMS], nh = 0, nout = 0;
    r = collide_and_collect(ctx, n, 0, &nh); if (r) return r;
    r = finalize_hits(ctx, nh, 0, true, &nout); if (r) return r;
    return deliver(ctx, nout, out, cap, count);
}

// ------------------------------------------------------------------------------------------------ broad phase
static bool fast_path_ok(const stf_ctx* ctx, int64_t nl) {
    int key_bits = 2 * STF_CELL_W + 5 + ctx->scene_bits;
    return !ctx->force_big && nl <= BS_CAP && key_bits + ctx->label_bits <= 64 && 2 * ctx->label_bits + FIN_IDX_BITS <= 64 && !ctx->env_force_big;
}
// locate + sort; on return *keys/*vals point at the sorted records, their number is in counts[C_NVALID] (device);
// *nrec_out = capacity bound (4 slots per located object).  `labels_dev` = the device copy of `labels` (or NULL).
// the sorts of the latency path: ONE CTA for small scenes (<= 2048 objects: <= 8192 records, and hit lists that practically
// never exceed the 5632 the single-CTA finalize holds — if one does, the round is re-run on the cluster kernels and the
// context stays there), the 8-CTA cluster otherwise (C1: records 28 vs 53 us, finalize 37 vs 45 us)
static inline bool single_cta_sorts(const stf_ctx* ctx) { return ctx->no_cluster || (ctx->n <= 2048 && !ctx->small_off); }
// fast: records appended compactly and sorted by ONE CTA (status bit STF_ST_RETRY_BIG if they do not fit);
// else: 4 fixed slots per object, multi-launch LSD radix sort, INVALID keys last.
static int32_t locate_and_sort(stf_ctx* ctx, int32_t nl, const int32_t* dl, int linked, bool fast, int64_t* nrec_out, uint64_t** keys, uint32_t** vals) {
    if (!ctx->sizelevel_ok) FAIL(STF_E_STATE, "locate!: an object is larger than the canvas (sizelevel == -1)");
    int64_t nrec = 4ll * (linked ? ctx->n : nl);
    *nrec_out = nrec;
    stage_begin(ctx, STF_STAGE_LOCATE);
    size_t kb = sizeof(uint64_t) * (size_t)std::max<int64_t>(4, nrec), vb = sizeof(uint32_t) * (size_t)std::max<int64_t>(4, nrec);
    CK(ctx->rk.ensure(kb)); CK(ctx->rk2.ensure(kb)); CK(ctx->rv.ensure(vb)); CK(ctx->rv2.ensure(vb));
    int bits = 2 * STF_CELL_W + 5 + ctx->scene_bits;
    unsigned long long* nvalid = ctx->counts.as<unsigned long long>() + C_NVALID;
    unsigned long long* nappend = ctx->counts.as<unsigned long long>() + C_NREC;
    if (fast) {
        const RoundCtl cancel = { ctx->in_rounds ? ctx->counts.as<unsigned long long>() + C_STOP : nullptr, ctx->dev_nl };
        if (linked) {
            if (nl > 0) LAUNCH_LOCATE(ctx, nl, ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), ctx->om.as<ObjMeta>(), ctx->L, nl, dl, 1,
                                                                   ctx->lkk.as<uint64_t>(), ctx->lkv.as<uint32_t>(), ERRP(ctx), nullptr, cancel);
            LAUNCH_PDL(ctx, k_compact_records, nblk(nrec, 256), 256, 0, ctx->lkk.as<uint64_t>(), ctx->lkv.as<uint32_t>(), nrec, ctx->rk2.as<uint64_t>(), ctx->rv2.as<uint32_t>(), nappend, cancel);
        } else if (nl > 0) {
            LAUNCH_LOCATE(ctx, nl, ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), ctx->om.as<ObjMeta>(), ctx->L, nl, dl, 0,
                                                        ctx->rk2.as<uint64_t>(), ctx->rv2.as<uint32_t>(), ERRP(ctx), nappend, cancel);
        }
        KCHECK();
        stage_begin(ctx, STF_STAGE_SORT);
        if (single_cta_sorts(ctx)) // small scenes: one CTA beats the cluster (no cluster barriers; measured on C2: 24 -> 20 us per round)
            LAUNCH_PDL(ctx, k_sort_records, 1, BS_THREADS, BS_SMEM_BYTES, ctx->rk2.as<uint64_t>(), ctx->rv2.as<uint32_t>(), nappend, nrec, bits, ctx->label_bits,
                                                                                      linked ? nullptr : dl, ctx->rk.as<uint64_t>(), ctx->rv.as<uint32_t>(), nvalid, STATUSP(ctx));
        else // one cluster of 8 CTAs
            LAUNCH_PDL(ctx, k_sort_records_cl, CS_CTAS, CS_THREADS, CS_SMEM_BYTES, ctx->rk2.as<uint64_t>(), ctx->rv2.as<uint32_t>(), nappend, nrec, bits, ctx->label_bits,
                                                                                               linked ? nullptr : dl, ctx->rk.as<uint64_t>(), ctx->rv.as<uint32_t>(), nvalid, STATUSP(ctx));
        KCHECK();
        *keys = ctx->rk.as<uint64_t>(); *vals = ctx->rv.as<uint32_t>();
        stage_end(ctx);
        return STF_OK; // ctr.records: from counts[C_NVALID] at the round's one synchronisation (read_round_counters)
    }
    CK(ctx->hist.ensure(sizeof(uint32_t) * radix_sort_hist_words(std::max<int64_t>(1, nrec))));
    if (linked) {
        if (nl > 0) LAUNCH_LOCATE(ctx, nl, ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), ctx->om.as<ObjMeta>(), ctx->L, nl, dl, 1,
                                                               ctx->lkk.as<uint64_t>(), ctx->lkv.as<uint32_t>(), ERRP(ctx), nullptr, RoundCtl{ nullptr, nullptr });
        CK(cudaMemcpyAsync(ctx->rk.p, ctx->lkk.p, sizeof(uint64_t) * (size_t)nrec, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaMemcpyAsync(ctx->rv.p, ctx->lkv.p, sizeof(uint32_t) * (size_t)nrec, cudaMemcpyDeviceToDevice, ctx->st));
    } else if (nl > 0) {
        LAUNCH_LOCATE(ctx, nl, ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), ctx->om.as<ObjMeta>(), ctx->L, nl, dl, 0,
                                                    ctx->rk.as<uint64_t>(), ctx->rv.as<uint32_t>(), ERRP(ctx), nullptr, RoundCtl{ nullptr, nullptr });
    }
    KCHECK();
    stage_begin(ctx, STF_STAGE_SORT);
    { // order-preserving compaction of the slot table (rk,rv) -> (rk2,rv2), then a stable sort over the varying key bits
        int nb = nblk(nrec, CMP_BLOCK);
        CK(ctx->keep.ensure(sizeof(uint32_t) * (size_t)(nb + 16)));
        unsigned long long* vary = ctx->counts.as<unsigned long long>() + C_VARY;
        CK(cudaMemsetAsync(nvalid, 0, sizeof(unsigned long long), ctx->st)); CK(cudaMemsetAsync(vary, 0, sizeof(unsigned long long), ctx->st));
        LAUNCH_PDL(ctx, k_compact_count, nb, CMP_BLOCK, 0, ctx->rk.as<uint64_t>(), nrec, ctx->keep.as<uint32_t>());
        LAUNCH_PDL(ctx, k_scan_u32, 1, 1024, 0, ctx->keep.as<uint32_t>(), nb, reinterpret_cast<uint32_t*>(nvalid));
        LAUNCH_PDL(ctx, k_compact_write, nb, CMP_BLOCK, 0, ctx->rk.as<uint64_t>(), ctx->rv.as<uint32_t>(), nrec, ctx->keep.as<uint32_t>(), ctx->rk2.as<uint64_t>(), ctx->rv2.as<uint32_t>());
        LAUNCH_PDL(ctx, k_vary_or, std::min(nblk(nrec / 4 + 1, 256), 148 * 4), 256, 0, ctx->rk2.as<uint64_t>(), nvalid, nrec, vary);
        KCHECK();
        int32_t r = sync_counts(ctx); if (r) return r;
        int64_t nv = (int64_t)ctx->h_counts[C_NVALID];
        ctx->ctr.records = nv; // valid (cell, label) records
        int f = radix_sort_pairs(ctx->st, &ctx->launches, ctx->rk2.as<uint64_t>(), ctx->rv2.as<uint32_t>(), ctx->rk.as<uint64_t>(), ctx->rv.as<uint32_t>(), nv,
                                 (uint64_t)ctx->h_counts[C_VARY] & bit_range(0, bits), ctx->hist.as<uint32_t>(), ctx->pdl, &ctx->after_kernel, ctx->sort3);
        *keys = f ? ctx->rk.as<uint64_t>() : ctx->rk2.as<uint64_t>();
        *vals = f ? ctx->rv.as<uint32_t>() : ctx->rv2.as<uint32_t>();
    }
    KCHECK();
    stage_end(ctx);
    return STF_OK;
}
static int32_t export_cells(stf_ctx* ctx, int64_t nrec, const uint64_t* keys, const uint32_t* vals, stf_cellrec* out, int64_t cap, int64_t* count) {
    CK(ctx->out32.ensure(sizeof(int32_t) * 4 * (size_t)std::max<int64_t>(1, nrec)));
    if (nrec > 0) ctx->launches++, k_export_cells<<<nblk(nrec, 256), 256, 0, ctx->st>>>(keys, vals, ctx->counts.as<unsigned long long>() + C_NVALID, nrec, ctx->out32.as<int32_t>());
    KCHECK();
    int32_t r = sync_counts(ctx); if (r) return r;
    int64_t nv = nrec > 0 ? (int64_t)ctx->h_counts[C_NVALID] : 0;
    if (count) *count = nv;
    if (nv > cap) FAIL(STF_E_CAPACITY, "output buffer too small");
    if (nv > 0) { CK(cudaMemcpyAsync(out, ctx->out32.p, sizeof(stf_cellrec) * (size_t)nv, cudaMemcpyDeviceToHost, ctx->st)); CK(cudaStreamSynchronize(ctx->st)); }
    return STF_OK;
}
extern "C" int32_t stf_locate(stf_ctx* ctx, int32_t nl, const int32_t* labels, stf_cellrec* out, int64_t cap, int64_t* count) {
    NEED_OBJS();
    if (!labels) nl = ctx->n;
    if (count) *count = 0;
    if (nl <= 0) return STF_OK;
    const int32_t* dl; int32_t r = upload_labels(ctx, ctx->lab, nl, labels, &dl); if (r) return r;
    int64_t nrec; uint64_t* keys; uint32_t* vals;
    r = locate_and_sort(ctx, nl, dl, 0, false, &nrec, &keys, &vals); if (r) return r;
    return export_cells(ctx, nrec, keys, vals, out, cap, count);
}
extern "C" int32_t stf_linked_reset(stf_ctx* ctx) {
    NEED_OBJS();
    if (ctx->n) ctx->launches++, k_fill_invalid<<<nblk(4ll * ctx->n, 256), 256, 0, ctx->st>>>(ctx->lkk.as<uint64_t>(), ctx->lkv.as<uint32_t>(), 4ll * ctx->n);
    KCHECK();
    CK(cudaStreamSynchronize(ctx->st));
    return STF_OK;
}
extern "C" int32_t stf_linked_locate(stf_ctx* ctx, int32_t nl, const int32_t* labels) {
    NEED_OBJS();
    if (!ctx->sizelevel_ok) FAIL(STF_E_STATE, "locate!: an object is larger than the canvas (sizelevel == -1)");
    if (!labels) nl = ctx->n;
    if (nl <= 0) return STF_OK;
    const int32_t* dl; int32_t r = upload_labels(ctx, ctx->lab, nl, labels, &dl); if (r) return r;
    LAUNCH_LOCATE(ctx, nl, ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), ctx->om.as<ObjMeta>(), ctx->L, nl, dl, 1, ctx->lkk.as<uint64_t>(),
                                                ctx->lkv.as<uint32_t>(), ERRP(ctx), nullptr, RoundCtl{ nullptr, nullptr });
    KCHECK();
    return sync_counts(ctx);
}
extern "C" int32_t stf_linked_clear(stf_ctx* ctx, int32_t nl, const int32_t* labels) {
    NEED_OBJS();
    if (nl <= 0 || !labels) return STF_OK;
    const int32_t* dl; int32_t r = upload_labels(ctx, ctx->lab, nl, labels, &dl); if (r) return r;
    ctx->launches++, k_clear_slots<<<nblk(nl, 128), 128, 0, ctx->st>>>(ctx->lkk.as<uint64_t>(), nl, dl);
    KCHECK();
    CK(cudaStreamSynchronize(ctx->st));
    return STF_OK;
}
extern "C" int32_t stf_linked_collect(stf_ctx* ctx, stf_cellrec* out, int64_t cap, int64_t* count) {
    NEED_OBJS();
    if (count) *count = 0;
    if (ctx->n == 0) return STF_OK;
    int64_t nrec; uint64_t* keys; uint32_t* vals;
    int32_t r = locate_and_sort(ctx, 0, nullptr, 1, false, &nrec, &keys, &vals); if (r) return r;
    return export_cells(ctx, nrec, keys, vals, out, cap, count);
}

// candidate generation.  fused: the narrow phase runs inside (decided pairs never become items; hits go to ctx->hits,
// undecided pairs to ctx->surv for launch_frontier); else the candidate list is materialised in ctx->items.
static int32_t launch_emit(stf_ctx* ctx, int64_t nrec, const uint64_t* keys, const uint32_t* vals, const int32_t* moved_pos, bool fused) {
    const int32_t* lvl_index = nullptr;
    if ((ctx->multi_scene || nrec > 65536) && nrec > 0) { // batches: (scene, level) index of the sorted records for the ancestor searches
        int nsl = 32 << ctx->scene_bits;
        CK(ctx->lvlidx.ensure(sizeof(int32_t) * (size_t)(nsl + 2)));
        LAUNCH_PDL(ctx, k_level_index, nblk(nrec + 1, 256), 256, 0, keys, ctx->counts.as<unsigned long long>() + C_NVALID, nrec, ctx->lvlidx.as<int32_t>(), nsl);
        lvl_index = ctx->lvlidx.as<int32_t>();
    }
    EmitOut eo; eo.items = ctx->items.as<Item16>(); eo.item_cap = (int64_t)(ctx->items.cap / sizeof(Item16));
    eo.direct = ctx->hits.as<Item16>(); eo.direct_cap = (int64_t)(ctx->hits.cap / sizeof(Item16));
    eo.counts = ctx->counts.as<unsigned long long>(); eo.nhits = ctx->counts.as<unsigned long long>() + C_NHITS;
    eo.surv = ctx->surv.as<SurvRec>(); eo.surv_cap = (int64_t)(ctx->surv.cap / sizeof(SurvRec)); eo.nsurv = ctx->counts.as<unsigned long long>() + C_SURV;
    eo.cursor = ctx->counts.as<unsigned long long>() + C_CURSOR;
    stage_begin(ctx, STF_STAGE_EMIT);
    if (nrec > 0) {
        int blocks = (int)std::min<int64_t>((nrec / 4 + EM_THREADS / 32 - 1) / (EM_THREADS / 32) + 1, 148 * 8); // ~1 record per located object
        // batches of scenes: few candidates per record and L <= 16 -> two records per warp
        const bool half = (ctx->multi_scene || nrec > 65536) && ctx->L <= 16;
#define STF_EMIT(W_, F_, B_) LAUNCH_PDL(ctx, (k_emit_items<W_, F_>), B_, EM_THREADS, 0, keys, vals, ctx->counts.as<unsigned long long>() + C_NVALID, nrec, ctx->L, ctx->words.as<uint32_t>(), \
                                        ctx->lm.as<LevelMeta>(), moved_pos, eo, ctx->shard_rank, ctx->shard_n, lvl_index)
        if (half) { if (fused) STF_EMIT(16, true, (blocks + 1) / 2); else STF_EMIT(16, false, (blocks + 1) / 2); }
        else { if (fused) STF_EMIT(32, true, blocks); else STF_EMIT(32, false, blocks); }
#undef STF_EMIT

    }
    KCHECK();
    stage_end(ctx);
    return STF_OK;
}
static int32_t upload_moved_pos(stf_ctx* ctx, int32_t nl, const int32_t* labels, const int32_t** moved_pos) {
    // moved_pos[o] = position of o+1 in `labels`, -1 otherwise
    std::vector<int32_t> mp((size_t)ctx->n, -1);
    for (int i = 0; i < nl; i++) mp[(size_t)(labels ? labels[i] - 1 : i)] = i;
    CK(stage_h2d(ctx, ctx->moved_pos.p, mp.data(), sizeof(int32_t) * (size_t)ctx->n));
    *moved_pos = ctx->moved_pos.as<int32_t>();
    return STF_OK;
}
static void read_round_counters(stf_ctx* ctx) {
    ctx->ctr.items = (int64_t)ctx->h_counts[C_ITEMS]; ctx->ctr.direct = (int64_t)ctx->h_counts[C_DIRECT]; ctx->ctr.hits = (int64_t)ctx->h_counts[C_NHITS];
    ctx->ctr.overflow = (int64_t)ctx->h_counts[C_OVERFLOW]; ctx->ctr.visited = (int64_t)ctx->h_counts[C_VISITED];
    ctx->ctr.records = (int64_t)ctx->h_counts[C_NVALID]; // valid (cell, label) records of the round's sort
    ctx->last_items = ctx->ctr.items; ctx->items_valid = !ctx->fused_narrow || ctx->keep_items; ctx->hint_items = ctx->ctr.items;
    ctx->hint_surv = (int64_t)ctx->h_counts[C_SURV];
}
// which list the finalize kernel checks for overflow before it lets the steps run: the candidate list when it is
// materialised, the survivor list of the fused narrow phase otherwise
static void set_overflow_check(stf_ctx* ctx, FinalizeArgs& F) {
    if (!ctx->fused_narrow || ctx->keep_items) { F.n_items_dev = ctx->counts.as<unsigned long long>() + C_ITEMS; F.item_cap = (int64_t)(ctx->items.cap / sizeof(Item16)); }
    else { F.n_items_dev = ctx->counts.as<unsigned long long>() + C_SURV; F.item_cap = (int64_t)(ctx->surv.cap / sizeof(SurvRec)); }
}

// Latency path of one many-body round: every count stays on the device, the host only enqueues.  Broad phase
// (append-locate, one-CTA sort, candidates) + narrow phase; the unordered hit list is left in ctx->hits with its count
// in counts[C_NHITS].  Buffer sizes come from the previous round; the finalize kernel raises RETRY bits if one was too
// small or the scene is too large for the single-CTA kernels, and the caller re-runs.
static int32_t spacial_core_fast(stf_ctx* ctx, int32_t nl, const int32_t* labels, int linked) {
    const int32_t* dl; int32_t r = STF_OK;
    if (ctx->dev_labels) dl = ctx->dev_labels; // the list was produced on the device; nl = its capacity, the count is read by k_locate
    else { r = upload_labels(ctx, ctx->lab, nl, labels, &dl); if (r) return r; }
    const int32_t* moved_pos = nullptr;
    if (linked) { r = upload_moved_pos(ctx, nl, labels, &moved_pos); if (r) return r; }
    int64_t nrec; uint64_t* keys; uint32_t* vals;
    if (!ctx->in_rounds) CK(cudaMemsetAsync(ctx->counts.p, 0, C_ERR * sizeof(unsigned long long), ctx->st)); // every counter and the status; pending error bits stay until a sync reads them
    // (inside stf_fit_rounds k_round_begin resets the counters and keeps the status: a raised bit cancels the rounds that follow)
    r = locate_and_sort(ctx, nl, dl, linked, true, &nrec, &keys, &vals); if (r) return r;
    CK(ctx->hits.ensure(sizeof(Item16) * (size_t)(BS_CAP + 1024))); // more hits than BS_CAP leave the fast path anyway
    if (ctx->fused_narrow && !ctx->keep_items) { // fused: candidates are tested where they are generated; only the undecided pairs are stored
        r = ensure_surv(ctx, std::max<int64_t>(2 * nrec / 4, ctx->hint_surv + ctx->hint_surv / 4 + 1024)); if (r) return r;
        r = launch_emit(ctx, nrec, keys, vals, moved_pos, true); if (r) return r;
        stage_begin(ctx, STF_STAGE_NARROW);
        NarrowArgs A; fill_narrow_args(ctx, A, false, true, true);
        r = launch_frontier(ctx, A, std::max<int64_t>(ctx->hint_surv, 512)); if (r) return r;
        stage_end(ctx);
        return STF_OK;
    }
    int64_t want_items = std::max<int64_t>(std::max<int64_t>(1024, 32 * nrec / 4), ctx->hint_items + ctx->hint_items / 4 + 1024);
    if ((int64_t)(ctx->items.cap / sizeof(Item16)) < want_items) CK(ctx->items.ensure(sizeof(Item16) * (size_t)want_items));
    r = launch_emit(ctx, nrec, keys, vals, moved_pos, false); if (r) return r;
    stage_begin(ctx, STF_STAGE_NARROW);
    r = narrow_phase(ctx, -1, (int64_t)(ctx->items.cap / sizeof(Item16)), false, true); if (r) return r;
    stage_end(ctx);
    return STF_OK;
}
// after the final sync of a fast round: 0 = done, 1 = re-run (buffers grown / path switched)
static int fast_round_verdict(stf_ctx* ctx) {
    int st = ctx->h_err[1];
    if (!st) return 0;
    cudaMemsetAsync(STATUSP(ctx), 0, sizeof(unsigned long long), ctx->st);
    if (st & STF_ST_RETRY_BIG) { if (single_cta_sorts(ctx) && !ctx->no_cluster) ctx->small_off = true; else ctx->force_big = true; } // (the cluster kernels hold more)
    if (st & STF_ST_RETRY_GROW) { ctx->hint_items = (int64_t)ctx->h_counts[C_ITEMS]; ctx->hint_surv = (int64_t)ctx->h_counts[C_SURV]; } // size the lists of the re-run
    return 1;
}

// throughput path: broad phase + narrow phase with host-known counts; leaves the raw hit list in ctx->hits, *nh hits
static int32_t spacial_core(stf_ctx* ctx, int32_t nl, const int32_t* labels, int linked, int64_t* nh) {
    *nh = 0;
    const int32_t* dl; int32_t r = upload_labels(ctx, ctx->lab, nl, labels, &dl); if (r) return r;
    const int32_t* moved_pos = nullptr;
    if (linked) { r = upload_moved_pos(ctx, nl, labels, &moved_pos); if (r) return r; }
    int64_t nrec; uint64_t* keys; uint32_t* vals;
    r = locate_and_sort(ctx, nl, dl, linked, false, &nrec, &keys, &vals); if (r) return r;
    int64_t n_items = 0, n_direct = 0;
    const bool fused = ctx->fused_narrow && !ctx->keep_items;
    unsigned long long* nsurv_dev = ctx->counts.as<unsigned long long>() + C_SURV;
    for (int attempt = 0; attempt < 3; attempt++) {
        r = zero_counts(ctx); if (r) return r;
        CK(cudaMemsetAsync(nsurv_dev, 0, sizeof(unsigned long long), ctx->st));
        if (ctx->items.cap < sizeof(Item16) * 1024 && !fused) CK(ctx->items.ensure(sizeof(Item16) * (size_t)std::max<int64_t>(1024, 32 * nrec)));
        if (ctx->hits.cap < sizeof(Item16) * 1024) CK(ctx->hits.ensure(sizeof(Item16) * (size_t)std::max<int64_t>(1024, nrec)));
        if (fused) { r = ensure_surv(ctx, std::max<int64_t>(nrec, ctx->hint_surv + ctx->hint_surv / 4 + 1024)); if (r) return r; }
        int64_t item_cap = (int64_t)(ctx->items.cap / sizeof(Item16)), hit_cap = (int64_t)(ctx->hits.cap / sizeof(Item16)), surv_cap = (int64_t)(ctx->surv.cap / sizeof(SurvRec));
        r = launch_emit(ctx, nrec, keys, vals, moved_pos, fused); if (r) return r;
        r = sync_counts(ctx); if (r) return r;
        n_items = (int64_t)ctx->h_counts[C_ITEMS]; n_direct = (int64_t)ctx->h_counts[C_DIRECT];
        const int64_t n_surv = (int64_t)ctx->h_counts[C_SURV], n_hits = (int64_t)ctx->h_counts[C_NHITS]; // fused: direct + narrow-phase hits so far
        if (fused ? (n_surv <= surv_cap && n_hits <= hit_cap) : (n_items <= item_cap && n_direct <= hit_cap)) break;
        if (attempt == 2) FAIL(STF_E_CUDA, "internal: candidate buffers did not converge");
        if (!fused) CK(ctx->items.ensure(sizeof(Item16) * (size_t)(n_items + n_items / 4 + 1024)));
        ctx->hint_surv = n_surv;
        CK(ctx->hits.ensure(sizeof(Item16) * (size_t)(n_hits + n_hits / 4 + 1024)));
    }
    stage_begin(ctx, STF_STAGE_NARROW);
    if (fused) { // the frontier kernels append at most one more hit per surviving pair
        const int64_t n_surv = (int64_t)ctx->h_counts[C_SURV], n_hits = (int64_t)ctx->h_counts[C_NHITS];
        CK(ctx->hits.ensure(sizeof(Item16) * (size_t)std::max<int64_t>(1, n_hits + n_surv), true, ctx->st));
        NarrowArgs A; fill_narrow_args(ctx, A, false, true, false);
        r = launch_frontier(ctx, A, n_surv); if (r) return r;
    } else {
        // direct hits are already at the front of ctx->hits (counts[C_NHITS] == n_direct); keep them while growing
        CK(ctx->hits.ensure(sizeof(Item16) * (size_t)std::max<int64_t>(1, n_items + n_direct), true, ctx->st));
        r = narrow_phase(ctx, n_items, n_items, false, true); if (r) return r;
    }
    stage_end(ctx);
    r = sync_counts(ctx); if (r) return r;
    read_round_counters(ctx);
    ctx->ctr.items = n_items; ctx->ctr.direct = n_direct; ctx->last_items = n_items;
    *nh = (int64_t)ctx->h_counts[C_NHITS];
    return STF_OK;
}

// many-body call with export: fast path when the scene fits, else the throughput path
static int32_t manybody_export(stf_ctx* ctx, int32_t nl, const int32_t* labels, int linked, int32_t unique, stf_item* out, int64_t cap, int64_t* count) {
    for (int attempt = 0; attempt < 4 && fast_path_ok(ctx, linked ? ctx->n : nl); attempt++) {
        int32_t r = spacial_core_fast(ctx, nl, labels, linked); if (r) return r;
        CK(ctx->hits2.ensure(sizeof(Item16) * (size_t)(BS_CAP + 1024))); CK(ctx->out32.ensure(sizeof(int32_t) * 5 * (size_t)(BS_CAP + 1024)));
        stage_begin(ctx, STF_STAGE_RESULT);
        FinalizeArgs F = {};
        F.hits = ctx->hits.as<Item16>(); F.nh_dev = ctx->counts.as<unsigned long long>() + C_NHITS; F.hits_cap = (int64_t)(ctx->hits.cap / sizeof(Item16));
        set_overflow_check(ctx, F);
        F.sorted = ctx->hits2.as<Item16>(); F.label_bits = ctx->label_bits; F.unique = unique;
        F.out5 = ctx->out32.as<int32_t>(); F.kept = ctx->counts.as<unsigned long long>() + C_KEPT; F.do_prep = 0; F.status = STATUSP(ctx);
        LAUNCH_PDL(ctx, k_finalize_hits, 1, BS_THREADS, BS_SMEM_BYTES, F);
        KCHECK();
        stage_end(ctx);
        r = sync_counts(ctx); if (r) return r;
        if (fast_round_verdict(ctx)) continue;
        read_round_counters(ctx);
        return deliver(ctx, (int64_t)ctx->h_counts[C_KEPT], out, cap, count);
    }
    int64_t nh = 0, nout = 0;
    int32_t r = spacial_core(ctx, nl, labels, linked, &nh); if (r) return r;
    r = finalize_hits(ctx, nh, unique, true, &nout); if (r) return r;
    return deliver(ctx, nout, out, cap, count);
}

extern "C" int32_t stf_totalcollisions_spacial(stf_ctx* ctx, int32_t nl, const int32_t* labels, int32_t unique, stf_item* out, int64_t cap, int64_t* count) {
    NEED_OBJS();
    if (count) *count = 0;
    ctx->last_result = 0;
    if (!labels) nl = ctx->n;
    if (ctx->n <= 1 || nl <= 0) return STF_OK; // length(qtrees) > 1 || return colist  :227
    return manybody_export(ctx, nl, labels, 0, unique, out, cap, count);
}
extern "C" int32_t stf_partialcollisions(stf_ctx* ctx, int32_t nl, const int32_t* labels, int32_t unique, stf_item* out, int64_t cap, int64_t* count) {
    NEED_OBJS();
    if (count) *count = 0;
    ctx->last_result = 0;
    if (!labels) nl = ctx->n;
    if (ctx->n == 0) return STF_OK;
    return manybody_export(ctx, nl, labels, 1, unique, out, cap, count);
}

// ------------------------------------------------------------------------------------------------ fit step
extern "C" int32_t stf_set_optimiser(stf_ctx* ctx, int32_t kind, double eta, double rho) {
    if (!ctx || kind < 0 || kind > 2) return STF_E_ARG;
    ctx->opt.kind = kind; ctx->opt.eta = eta; ctx->opt.rho = rho;
    return STF_OK;
}
extern "C" int32_t stf_reset_velocity(stf_ctx* ctx, int32_t n, const int32_t* labels) {
    NEED_OBJS();
    if (!labels) { if (ctx->n) CK(cudaMemsetAsync(ctx->has_vel.p, 0, (size_t)ctx->n, ctx->st)); }
    else for (int i = 0; i < n; i++) { CHECK_LABEL(labels[i]); CK(cudaMemsetAsync(ctx->has_vel.as<uint8_t>() + (labels[i] - 1), 0, 1, ctx->st)); }
    return STF_OK;
}
extern "C" int32_t stf_get_velocity(stf_ctx* ctx, int32_t n, const int32_t* labels, double* v, uint8_t* has) {
    NEED_OBJS();
    std::vector<double> hv(2 * (size_t)std::max(1, ctx->n)); std::vector<uint8_t> hh((size_t)std::max(1, ctx->n));
    if (ctx->n) {
        CK(cudaMemcpyAsync(hv.data(), ctx->vel.p, sizeof(double) * 2 * (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(hh.data(), ctx->has_vel.p, (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->st));
    }
    CK(cudaStreamSynchronize(ctx->st));
    for (int i = 0; i < n; i++) {
        int lb = labels ? labels[i] : i + 1; CHECK_LABEL(lb);
        has[i] = hh[lb - 1]; v[2 * i] = has[i] ? hv[2 * (size_t)(lb - 1)] : 0.0; v[2 * i + 1] = has[i] ? hv[2 * (size_t)(lb - 1) + 1] : 0.0;
    }
    return STF_OK;
}
// k_batchsteps + k_materialize over `hits`; the step count is host-known (n_dev == NULL) or lives on the device
static int32_t launch_steps(stf_ctx* ctx, const Item16* hits, int64_t n_host, const unsigned long long* n_dev, uint64_t seed, uint64_t iter, const uint32_t* order = nullptr, bool by_scene = false) {
    StepArgs A;
    A.words = ctx->words.as<uint32_t>(); A.lm = ctx->lm.as<LevelMeta>(); A.om = ctx->om.as<ObjMeta>(); A.L = ctx->L;
    A.hits = hits; A.n = n_host; A.n_dev = n_dev; A.prev = ctx->prev.as<uint32_t>(); A.done = ctx->done.as<uint32_t>();
    A.ticket = ctx->counts.as<unsigned long long>() + C_TICKET; A.order = order;
    A.vel = ctx->vel.as<double>(); A.has_vel = ctx->has_vel.as<uint8_t>(); A.moved_flag = ctx->moved_flag.as<uint8_t>(); A.dirty = ctx->dirty.as<int>(); A.built_from = ctx->built_from.as<int>();
    A.opt = ctx->opt; A.seed = seed; A.iter = iter; A.counts = ctx->counts.as<unsigned long long>();
    A.dbg = ctx->dbg_step;
    A.dbg_t = nullptr;
    if (A.dbg & 16) { CK(ctx->dbg_t.ensure(sizeof(unsigned long long) * 4 * (size_t)std::max<int64_t>(n_host, BS_CAP))); A.dbg_t = ctx->dbg_t.as<unsigned long long>(); }
    int blocks = (int)std::min<int64_t>(std::max<int64_t>(1, (n_host + 3) / 4), 148 * 8);
    stage_begin(ctx, STF_STAGE_STEP);
#ifndef STF_SCENE_NT
#define STF_SCENE_NT 128
#define STF_SCENE_MINB 6
#endif
#define K_STEPS_SCENE k_batchsteps_scene<STF_SCENE_NT, 256, STF_SCENE_MINB>
    if (by_scene) LAUNCH_PDL(ctx, K_STEPS_SCENE, std::min(ctx->n_scenes, 148 * STF_SCENE_MINB), STF_SCENE_NT, 0, A, (const uint32_t*)ctx->scene_rng.as<uint32_t>(), ctx->n_scenes);
    else if (n_host > 32768) LAUNCH_PDL(ctx, k_batchsteps<6>, blocks, 128, 0, A);
    else LAUNCH_PDL(ctx, k_batchsteps<4>, blocks, 128, 0, A);
    // the steps only moved the shifts: materialise the rasters of every moved tree once
    { int grp = rebuild_group(ctx->n);
      LAUNCH_PDL(ctx, k_materialize, std::min(nblk(((int64_t)ctx->n + grp - 1) / grp * 32, 256), 148 * 8), 256, 0, ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), ctx->L, ctx->n, ctx->dirty.as<int>(), ctx->built_from.as<int>(), grp); }
    KCHECK();
    stage_end(ctx);
    return STF_OK;
}
// batchsteps! over hits[0..n) (device, list order, host-known n)
// scenes_contiguous: the steps of one scene are consecutive in the list (true for the canonical order of a batch whose scene
// ids ascend with the labels; a caller's own list is checked on the host): the precondition of the scene-by-scene step kernel
static int32_t batchsteps_device(stf_ctx* ctx, const Item16* hits, int64_t n, uint64_t seed, uint64_t iter, bool scenes_contiguous) {
    if (n == 0) { ctx->ctr.moved = 0; return STF_OK; }
    int64_t m = 2 * n;
    CK(ctx->prev.ensure(sizeof(uint32_t) * (size_t)m)); CK(ctx->done.ensure(sizeof(uint32_t) * (size_t)n));
    const uint32_t* order = nullptr; bool by_scene = false;
    stage_begin(ctx, STF_STAGE_STEP_PREP);
    CK(cudaMemsetAsync(ctx->counts.as<unsigned long long>() + C_MOVES, 0, sizeof(unsigned long long), ctx->st));
    if (m <= BS_CAP && ctx->label_bits + 1 + FIN_IDX_BITS <= 64) { // one CTA: endpoints, sort, links
        LAUNCH_PDL(ctx, k_step_links, 1, BS_THREADS, BS_SMEM_BYTES, hits, (int)n, ctx->label_bits, ctx->om.as<ObjMeta>(), ctx->prev.as<uint32_t>(), ctx->done.as<uint32_t>(),
                                                                                 ctx->counts.as<unsigned long long>() + C_TICKET);
    } else {
        CK(ctx->hk.ensure(sizeof(uint64_t) * (size_t)m)); CK(ctx->hk2.ensure(sizeof(uint64_t) * (size_t)m));
        CK(ctx->hv.ensure(sizeof(uint32_t) * (size_t)m)); CK(ctx->hv2.ensure(sizeof(uint32_t) * (size_t)m));
        CK(ctx->hist.ensure(sizeof(uint32_t) * radix_sort_hist_words(m)));
        LAUNCH_PDL(ctx, k_step_endpoints, nblk(m, 256), 256, 0, hits, n, ctx->om.as<ObjMeta>(), ctx->hk.as<uint64_t>(), ctx->hv.as<uint32_t>());
        CK(cudaMemsetAsync(ctx->prev.p, 0, sizeof(uint32_t) * (size_t)m, ctx->st));
        int f = radix_sort_pairs(ctx->st, &ctx->launches, ctx->hk.as<uint64_t>(), ctx->hv.as<uint32_t>(), ctx->hk2.as<uint64_t>(), ctx->hv2.as<uint32_t>(), m, bit_range(0, ctx->label_bits + 1), ctx->hist.as<uint32_t>(), ctx->pdl, &ctx->after_kernel, ctx->sort3);
        CK(cudaMemsetAsync(ctx->done.p, 0, sizeof(uint32_t) * (size_t)n, ctx->st));
        CK(cudaMemsetAsync(ctx->counts.as<unsigned long long>() + C_TICKET, 0, sizeof(unsigned long long), ctx->st));
        LAUNCH_PDL(ctx, k_step_prev, nblk(m, 256), 256, 0, f ? ctx->hk2.as<uint64_t>() : ctx->hk.as<uint64_t>(), f ? ctx->hv2.as<uint32_t>() : ctx->hv.as<uint32_t>(), nullptr, m, ctx->prev.as<uint32_t>());
        if (n >= 16384 && ctx->multi_scene && scenes_contiguous && !ctx->step_global) { // a batch of scenes: one CTA per scene, levels separated by CTA barriers
            CK(ctx->scene_rng.ensure(sizeof(uint32_t) * 2 * (size_t)ctx->n_scenes));
            CK(cudaMemsetAsync(ctx->scene_rng.p, 0, sizeof(uint32_t) * 2 * (size_t)ctx->n_scenes, ctx->st));
            LAUNCH_PDL(ctx, k_scene_ranges, nblk(n, 256), 256, 0, hits, n, ctx->om.as<ObjMeta>(), ctx->scene_rng.as<uint32_t>());
            by_scene = true;
        } else if (n >= 16384) { // many chains (a batch of scenes): claim the steps level by level
            CK(ctx->keep.ensure(sizeof(uint32_t) * 2 * (size_t)n));
            uint32_t* la = ctx->keep.as<uint32_t>(); uint32_t* lb = la + n;
            int level_bits = 5;
            if (n <= 32768 && !ctx->sort3) { // one launch: exact depths (clamped to 63) by one CTA
                LAUNCH_PDL(ctx, k_step_depth_cta, 1, 1024, 0, (const uint32_t*)ctx->prev.as<uint32_t>(), la, n);
                level_bits = 6;
            } else {
                CK(cudaMemsetAsync(la, 0, sizeof(uint32_t) * (size_t)n, ctx->st));
                for (int sweep = 0; sweep < 16; sweep++) { LAUNCH_PDL(ctx, k_step_level, nblk(n, 256), 256, 0, ctx->prev.as<uint32_t>(), la, lb, n); std::swap(la, lb); }
            }
            LAUNCH_PDL(ctx, k_step_level_keys, nblk(n, 256), 256, 0, la, n, ctx->hk.as<uint64_t>(), ctx->hv.as<uint32_t>());
            int f2 = radix_sort_pairs(ctx->st, &ctx->launches, ctx->hk.as<uint64_t>(), ctx->hv.as<uint32_t>(), ctx->hk2.as<uint64_t>(), ctx->hv2.as<uint32_t>(), n, bit_range(0, level_bits), ctx->hist.as<uint32_t>(), ctx->pdl, &ctx->after_kernel, ctx->sort3);
            order = f2 ? ctx->hv2.as<uint32_t>() : ctx->hv.as<uint32_t>();
        }
    }
    KCHECK();
    int32_t r = launch_steps(ctx, hits, n, nullptr, seed, iter, order, by_scene); if (r) return r;
    r = sync_counts(ctx); if (r) return r;
    ctx->ctr.moved = (int64_t)ctx->h_counts[C_MOVES];
    return STF_OK;
}
extern "C" int32_t stf_batchsteps(stf_ctx* ctx, int64_t n, const stf_item* colist, uint64_t seed, uint64_t iteration) {
    NEED_OBJS();
    if (n <= 0) return STF_OK;
    bool contiguous = ctx->multi_scene; int last_scene = -1; // (a caller's list is stepped in ITS order: scene by scene only if it is grouped by scene)
    for (int64_t i = 0; i < n; i++) {
        if (colist[i].l < 1) FAIL(STF_E_ARG, "collision point level must be >= 1");
        if (contiguous) {
            const int32_t a = colist[i].i1, b = colist[i].i2;
            if (a < 1 || a > ctx->n || b < 1 || b > ctx->n) { contiguous = false; continue; } // (import_items reports the bad label)
            const int sa = ctx->h_om[a - 1].scene, sb = ctx->h_om[b - 1].scene;
            if (sa != sb || sa < last_scene) contiguous = false;
            last_scene = sa;
        }
    }
    int32_t r = import_items(ctx, n, colist, ctx->hits2); if (r) return r;
    return batchsteps_device(ctx, ctx->hits2.as<Item16>(), n, seed, iteration, contiguous);
}
// One fit iteration on the device: many-body collision detection over the active labels, then batchsteps! on the result
// (canonical list order = sorted by ((i1,i2),(l,a,b))).  Latency path: ~10 launches, ONE host synchronisation.
// positions of every tree at ctx->pos_level -> pinned landing area (enqueue only); pos_deliver copies them out after the sync
static int32_t pos_enqueue(stf_ctx* ctx) {
    const int n = ctx->n; const size_t nb = sizeof(int32_t) * (size_t)n;
    CK(ctx->out32.ensure(2 * nb));
    if (ctx->h_out_cap < 2 * nb) {
        if (ctx->h_out) cudaFreeHost(ctx->h_out);
        ctx->h_out = nullptr; ctx->h_out_cap = 0;
        CK(cudaMallocHost(&ctx->h_out, 4 * nb)); ctx->h_out_cap = 4 * nb;
    }
    if (ctx->zero_copy) { // posted writes over PCIe into the pinned landing area: no copy-engine operation before the sync
        LAUNCH_PDL(ctx, k_get_shifts, nblk(n, 256), 256, 0, ctx->lm.as<LevelMeta>(), ctx->L, n, nullptr, ctx->pos_level, reinterpret_cast<int32_t*>(ctx->h_out));
        KCHECK();
        return STF_OK;
    }
    LAUNCH_PDL(ctx, k_get_shifts, nblk(n, 256), 256, 0, ctx->lm.as<LevelMeta>(), ctx->L, n, nullptr, ctx->pos_level, ctx->out32.as<int32_t>());
    KCHECK();
    CK(cudaMemcpyAsync(ctx->h_out, ctx->out32.p, 2 * nb, cudaMemcpyDeviceToHost, ctx->st));
    return STF_OK;
}
static void pos_deliver(stf_ctx* ctx) {
    const size_t nb = sizeof(int32_t) * (size_t)ctx->n;
    memcpy(ctx->pos_rs, ctx->h_out, nb); memcpy(ctx->pos_cs, ctx->h_out + nb, nb);
}
// second half of a latency-path round: the unordered hit list (ctx->hits, count on the device) -> canonical order + step
// links (one cluster kernel) -> batchsteps + materialisation.  Enqueue only.  cancel_on_status: a status bit raised before
// (by this GPU, or by a peer of an item-parallel group) cancels the round.
static int32_t fast_round_tail(stf_ctx* ctx, uint64_t seed, uint64_t iteration, bool cancel_on_status) {
    CK(ctx->hits2.ensure(sizeof(Item16) * (size_t)(BS_CAP + 1024)));
    CK(ctx->prev.ensure(sizeof(uint32_t) * (size_t)(CS_CAP + 1024))); CK(ctx->done.ensure(sizeof(uint32_t) * (size_t)(BS_CAP + 1024))); // two links per step, up to CS_CAP / 2 steps (the cluster finalize)
    stage_begin(ctx, STF_STAGE_RESULT);
    FinalizeArgs F = {};
    F.hits = ctx->hits.as<Item16>(); F.nh_dev = ctx->counts.as<unsigned long long>() + C_NHITS; F.hits_cap = (int64_t)(ctx->hits.cap / sizeof(Item16));
    set_overflow_check(ctx, F);
    F.sorted = ctx->hits2.as<Item16>(); F.label_bits = ctx->label_bits; F.unique = 0; F.out5 = nullptr; F.kept = nullptr;
    F.do_prep = 1; F.om = ctx->om.as<ObjMeta>(); F.prev = ctx->prev.as<uint32_t>(); F.done = ctx->done.as<uint32_t>();
    F.ticket = ctx->counts.as<unsigned long long>() + C_TICKET; F.n_steps = ctx->counts.as<unsigned long long>() + C_NSTEPS;
    F.status = STATUSP(ctx); F.abort_on_status = cancel_on_status ? 1 : 0;
    F.stop = ctx->in_rounds ? ctx->counts.as<unsigned long long>() + C_STOP : nullptr;
    if (single_cta_sorts(ctx)) LAUNCH_PDL(ctx, k_finalize_hits, 1, BS_THREADS, BS_SMEM_BYTES, F);
    else {
        CK(ctx->hk.ensure(sizeof(uint64_t) * (size_t)(CS_CAP + 1024)));
        LAUNCH_PDL(ctx, k_finalize_prep_cl, CS_CTAS, CS_THREADS, CS_SMEM_BYTES, F, ctx->hk.as<uint64_t>());
    }
    KCHECK();
    return launch_steps(ctx, ctx->hits2.as<Item16>(), std::max<int64_t>(ctx->ctr.hits, 2048), ctx->counts.as<unsigned long long>() + C_NSTEPS, seed, iteration);
}
extern "C" int32_t stf_fit_round_positions(stf_ctx* ctx, int32_t mode, int32_t nl, const int32_t* labels, uint64_t seed, uint64_t iteration, int64_t* nhits,
                                           int32_t level, int32_t* rs, int32_t* cs) {
    NEED_OBJS();
    if (level < 1 || level > ctx->L || !rs || !cs) FAIL(STF_E_ARG, "bad level or output arrays");
    ctx->pos_level = level; ctx->pos_rs = rs; ctx->pos_cs = cs;
    int32_t r = stf_fit_round(ctx, mode, nl, labels, seed, iteration, nhits);
    bool delivered = ctx->pos_level == 0;
    ctx->pos_level = 0;
    if (r) return r;
    return delivered ? STF_OK : stf_get_shifts(ctx, ctx->n, nullptr, level, rs, cs); // paths without a folded read-back
}
// the round + the distinct labels of its colist: `moved := labels of the result` of dynamiccollisions (qtree_functions.jl:354-355)
// and `inds` of the element-wise trainers (fit.jl:90) without shipping the hit list to the host.  Ascending label order.
extern "C" int32_t stf_fit_round_moved(stf_ctx* ctx, int32_t mode, int32_t nl, const int32_t* labels, uint64_t seed, uint64_t iteration, int64_t* nhits,
                                       int32_t* moved_out, int32_t cap, int32_t* nmoved) {
    if (!ctx) return STF_E_ARG;
    if (nmoved) *nmoved = 0;
    const bool device_list = ctx->L > 0 && ctx->n <= 65536;
    ctx->want_moved = device_list;
    if (ctx->lab_mark.p && ctx->lab_mark.cap < sizeof(uint32_t) * (size_t)ctx->n) ctx->lab_mark.release();
    int64_t nh = 0;
    int32_t r = stf_fit_round(ctx, mode, nl, labels, seed, iteration, &nh);
    const bool folded = ctx->want_moved == false && device_list; // stf_fit_round clears the flag when the latency path delivered the list
    ctx->want_moved = false;
    if (r) return r;
    if (nhits) *nhits = nh;
    std::vector<int32_t> lab;
    if (folded) {
        const int32_t* src = reinterpret_cast<const int32_t*>(ctx->h_out);
        lab.assign(src + 1, src + 1 + src[0]);
    } else if (nh > 0) { // throughput path: the sorted hit list is still on the device (ctx->hits2)
        std::vector<Item16> h((size_t)nh);
        CK(cudaMemcpyAsync(h.data(), ctx->hits2.p, sizeof(Item16) * (size_t)nh, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        for (const Item16& x : h) { lab.push_back(x.i1); lab.push_back(item_i2(x)); }
        std::sort(lab.begin(), lab.end()); lab.erase(std::unique(lab.begin(), lab.end()), lab.end());
    }
    if (nmoved) *nmoved = (int32_t)lab.size();
    if ((int64_t)lab.size() > cap) FAIL(STF_E_CAPACITY, "label buffer too small");
    if (!lab.empty()) { if (!moved_out) FAIL(STF_E_ARG, "null output buffer"); memcpy(moved_out, lab.data(), sizeof(int32_t) * lab.size()); }
    return STF_OK;
}
extern "C" int32_t stf_fit_round(stf_ctx* ctx, int32_t mode, int32_t nl, const int32_t* labels, uint64_t seed, uint64_t iteration, int64_t* nhits) {
    NEED_OBJS();
    if (nhits) *nhits = 0;
    if (!labels) nl = ctx->n;
    if (ctx->n <= 1 || nl <= 0) return STF_OK;
    ctx->last_result = 0;
    int linked = mode == 1;
    for (int attempt = 0; attempt < 4 && fast_path_ok(ctx, linked ? ctx->n : nl); attempt++) {
        int32_t r = spacial_core_fast(ctx, nl, labels, linked); if (r) return r;
        r = fast_round_tail(ctx, seed, iteration, false); if (r) return r;
        if (ctx->want_moved) { // the labels of the colist (ascending), written straight into the pinned landing area
            const size_t need = sizeof(int32_t) * (size_t)(ctx->n + 1);
            if (ctx->h_out_cap < need) { if (ctx->h_out) cudaFreeHost(ctx->h_out); ctx->h_out = nullptr; ctx->h_out_cap = 0; CK(cudaMallocHost(&ctx->h_out, 2 * need)); ctx->h_out_cap = 2 * need; }
            if (!ctx->lab_mark.p) { CK(ctx->lab_mark.ensure(sizeof(uint32_t) * (size_t)ctx->n)); CK(cudaMemsetAsync(ctx->lab_mark.p, 0, sizeof(uint32_t) * (size_t)ctx->n, ctx->st)); ctx->lab_epoch = 0; }
            LAUNCH_PDL(ctx, k_result_labels, 1, 1024, 0, (const Item16*)ctx->hits2.as<Item16>(), (const unsigned long long*)(ctx->counts.as<unsigned long long>() + C_NSTEPS), (int64_t)BS_CAP,
                       ctx->lab_mark.as<uint32_t>(), ctx->n, ++ctx->lab_epoch, reinterpret_cast<int32_t*>(ctx->h_out), ctx->n + 1);
            KCHECK();
        }
        const bool fold = ctx->pos_level > 0 && ctx->n <= 65536 && !ctx->want_moved;
        if (fold) { r = pos_enqueue(ctx); if (r) return r; }
        r = sync_counts(ctx); if (r) return r;
        if (fast_round_verdict(ctx)) continue; // no step ran (n_steps == 0): nothing to undo
        if (fold) { pos_deliver(ctx); ctx->pos_level = 0; }
        ctx->want_moved = false; // (delivered: the caller reads the label list from the landing area)
        read_round_counters(ctx);
        ctx->ctr.moved = (int64_t)ctx->h_counts[C_MOVES];
        if (nhits) *nhits = ctx->ctr.hits;
        return STF_OK;
    }
    int64_t nh = 0, nout = 0;
    int32_t r = spacial_core(ctx, nl, labels, linked, &nh); if (r) return r;
    r = finalize_hits(ctx, nh, 0, false, &nout); if (r) return r; // canonical list order = sorted, stays on the device
    if (nhits) *nhits = nh;
    return batchsteps_device(ctx, ctx->hits2.as<Item16>(), nh, seed, iteration, ctx->scene_sorted);
}
// ---- several fit rounds without a host round trip in between (the inner loop of the trainers,
// fit.jl:85-99, stays on the device).  Every round is enqueued blindly; k_round_begin resets the round's counters on the
// device and cancels the round (C_STOP) when an earlier round raised a status bit (a list was too small / the scene left the
// latency path) — a cancelled round's locate produces no records, so all its kernels find empty lists; k_round_end records
// (hits, items, node tests, moves) of the round.  ONE synchronisation per call.
#define ROUND_REC 4
__global__ void k_round_begin(unsigned long long* __restrict__ counts) {
    pdl_sync();
    const int t = threadIdx.x;
    const int status = (int)counts[C_STATUS];
    const bool stop = status != 0 || counts[C_STOP] != 0ull;
    __syncwarp();
    if (t == 0 && stop && counts[C_STOP] == 0ull) counts[C_STOP] = 2ull; // a previous round failed: nothing of what follows may run (other stop codes stay)
    // a cancelled round only empties the lists (its kernels find nothing to do); the counters of the failed round stay: the
    // host sizes the buffers of the re-run from them
    // (an epoch that simply ended — no status bit — also drops the stale candidate list: nothing may be tested or stepped again)
    if (stop) { if (t == C_NREC || t == C_NVALID || t == C_NHITS || t == C_NSTEPS || t == C_SURV || (t == C_ITEMS && status == 0)) counts[t] = 0ull; return; }
    if (t < C_STATUS) counts[t] = 0ull;
}
__global__ void k_round_end(unsigned long long* __restrict__ counts, unsigned long long* __restrict__ rec, int k) {
    pdl_sync();
    if (threadIdx.x != 0) return;
    if (counts[C_STOP] != 0ull) return;                 // cancelled round: no record
    if ((int)counts[C_STATUS] != 0) { counts[C_STOP] = 2ull; return; } // this round raised a bit (its steps did not run)
    unsigned long long* r = rec + (size_t)ROUND_REC * (k + 1);
    r[0] = counts[C_NHITS]; r[1] = counts[C_ITEMS] + counts[C_DIRECT]; r[2] = counts[C_VISITED]; r[3] = counts[C_MOVES];
    rec[0] = (unsigned long long)(k + 1);               // rounds done
    rec[1] = counts[C_NVALID];
}
extern "C" int32_t stf_fit_rounds(stf_ctx* ctx, int32_t mode, int32_t nl, const int32_t* labels, uint64_t seed, uint64_t iteration0, int32_t n_rounds,
                                  int64_t* nhits_per_round, stf_counters_t* totals, int32_t level, int32_t* rs, int32_t* cs) {
    NEED_OBJS();
    if (n_rounds < 0 || (level != 0 && (level < 1 || level > ctx->L || !rs || !cs))) FAIL(STF_E_ARG, "bad round count, level or output arrays");
    if (!labels) nl = ctx->n;
    stf_counters_t tot = {};
    for (int k = 0; k < n_rounds; k++) if (nhits_per_round) nhits_per_round[k] = 0;
    const int linked = mode == 1;
    int done = 0;
    if (ctx->n > 1 && nl > 0) {
        CK(ctx->rounds_rec.ensure(sizeof(unsigned long long) * ROUND_REC * (size_t)(n_rounds + 2)));
        std::vector<unsigned long long> rec((size_t)ROUND_REC * (n_rounds + 2));
        while (done < n_rounds) {
            if (!fast_path_ok(ctx, linked ? ctx->n : nl)) { // scenes beyond the latency path: one host-driven round at a time
                int64_t nh = 0;
                int32_t r = stf_fit_round(ctx, mode, nl, labels, seed, iteration0 + done, &nh); if (r) return r;
                if (nhits_per_round) nhits_per_round[done] = nh;
                tot.hits += nh; tot.items += ctx->ctr.items; tot.direct += ctx->ctr.direct; tot.visited += ctx->ctr.visited; tot.moved += nh ? ctx->ctr.moved : 0; tot.records = ctx->ctr.records;
                done++;
                continue;
            }
            const int todo = n_rounds - done;
            CK(cudaMemsetAsync(ctx->rounds_rec.p, 0, sizeof(unsigned long long) * ROUND_REC * (size_t)(todo + 2), ctx->st));
            CK(cudaMemsetAsync(ctx->counts.as<unsigned long long>() + C_STATUS, 0, sizeof(unsigned long long), ctx->st));
            CK(cudaMemsetAsync(ctx->counts.as<unsigned long long>() + C_STOP, 0, sizeof(unsigned long long), ctx->st));
            ctx->in_rounds = true;
            int32_t r = STF_OK;
            for (int k = 0; k < todo && !r; k++) {
                LAUNCH_PDL(ctx, k_round_begin, 1, 32, 0, ctx->counts.as<unsigned long long>());
                r = spacial_core_fast(ctx, nl, labels, linked);
                if (!r) r = fast_round_tail(ctx, seed, iteration0 + done + k, true);
                if (!r) LAUNCH_PDL(ctx, k_round_end, 1, 32, 0, ctx->counts.as<unsigned long long>(), ctx->rounds_rec.as<unsigned long long>(), k);
            }
            ctx->in_rounds = false;
            if (r) return r;
            const bool fold = level > 0 && ctx->n <= 65536;
            if (fold) { ctx->pos_level = level; ctx->pos_rs = rs; ctx->pos_cs = cs; r = pos_enqueue(ctx); if (r) { ctx->pos_level = 0; return r; } }
            CK(cudaMemcpyAsync(rec.data(), ctx->rounds_rec.p, sizeof(unsigned long long) * ROUND_REC * (size_t)(todo + 2), cudaMemcpyDeviceToHost, ctx->st));
            r = sync_counts(ctx);
            ctx->pos_level = 0;
            if (r) return r;
            const int ran = (int)rec[0];
            for (int k = 0; k < ran; k++) {
                const unsigned long long* q = rec.data() + (size_t)ROUND_REC * (k + 1);
                if (nhits_per_round) nhits_per_round[done + k] = (int64_t)q[0];
                tot.hits += (int64_t)q[0]; tot.items += (int64_t)q[1]; tot.visited += (int64_t)q[2]; tot.moved += q[0] ? (int64_t)q[3] : 0;
            }
            tot.records = (int64_t)rec[1];
            done += ran;
            CK(cudaMemsetAsync(ctx->counts.as<unsigned long long>() + C_STOP, 0, sizeof(unsigned long long), ctx->st));
            if (ran < todo) { // round `done` raised a status bit: grow / switch path, then go on from there
                if (!fast_round_verdict(ctx)) FAIL(STF_E_CUDA, "internal: a round was cancelled without a status bit");
                continue;
            }
            if (fold) { pos_deliver(ctx); level = 0; }
        }
    } else done = n_rounds;
    ctx->ctr.hits = nhits_per_round && n_rounds > 0 ? nhits_per_round[n_rounds - 1] : 0;
    if (totals) *totals = tot;
    if (level > 0) return stf_get_shifts(ctx, ctx->n, nullptr, level, rs, cs); // paths without a folded read-back
    return STF_OK;
}

// trainepoch_E! (fit.jl:85-99) with its round loop on the device: round 0 over every label, `inds` (the labels of its
// colist, first-appearance order) and the round limit built on the device, then rounds over `inds` until the reference's
// stop rule fires — all enqueued blindly, ONE synchronisation per call.  first_round = 0 starts an epoch; > 0 resumes the
// epoch whose state is still on the device at that round (after max_rounds ran out).  state_out = {reason, ni, ninds,
// limit}: reason 0 = max_rounds ran out (call again with first_round = rounds done so far), 1 = the epoch is finished,
// 3 = `inds` (returned in inds_out) is at or below SPACIAL_ENABLE_THRESHOLD = 10: the dispatcher's native path takes over
// and the caller runs the remaining rounds ni+1.. through totalcollisions(qtrees, inds) + batchsteps! itself.
extern "C" int32_t stf_fit_epoch_e(stf_ctx* ctx, uint64_t seed, uint64_t iteration0, int32_t first_round, int32_t max_rounds,
                                   int64_t* nhits_per_round, int32_t* rounds_done, int32_t state_out[4], int32_t* inds_out, int32_t inds_cap) {
    NEED_OBJS();
    if (rounds_done) *rounds_done = 0;
    if (max_rounds < 1 || first_round < 0 || !state_out) FAIL(STF_E_ARG, "bad round counts or null state");
    if (ctx->n <= 1) { state_out[0] = 1; state_out[1] = state_out[2] = state_out[3] = 0; return STF_OK; }
    if (!fast_path_ok(ctx, ctx->n)) FAIL(STF_E_STATE, "stf_fit_epoch_e needs a scene on the single-scene latency path (use the host trainer over stf_fit_round)");
    CK(ctx->rounds_rec.ensure(sizeof(unsigned long long) * ROUND_REC * (size_t)(max_rounds + 2)));
    CK(ctx->inds.ensure(sizeof(int32_t) * (size_t)ctx->n)); CK(ctx->epoch_ctl.ensure(sizeof(int) * 8));
    if (ctx->lab_mark.cap < sizeof(uint32_t) * (size_t)ctx->n) { ctx->lab_mark.release(); CK(ctx->lab_mark.ensure(sizeof(uint32_t) * (size_t)ctx->n)); CK(cudaMemsetAsync(ctx->lab_mark.p, 0, sizeof(uint32_t) * (size_t)ctx->n, ctx->st)); ctx->lab_epoch = 0; }
    std::vector<unsigned long long> rec((size_t)ROUND_REC * (max_rounds + 2));
    int ctl[4] = {};
    int done = 0, round = first_round;
    unsigned long long* C0 = ctx->counts.as<unsigned long long>();
    for (int attempt = 0; attempt < 6; attempt++) {
        const int todo = max_rounds - done;
        CK(cudaMemsetAsync(ctx->rounds_rec.p, 0, sizeof(unsigned long long) * ROUND_REC * (size_t)(todo + 2), ctx->st));
        CK(cudaMemsetAsync(C0 + C_STATUS, 0, sizeof(unsigned long long), ctx->st));
        CK(cudaMemsetAsync(C0 + C_STOP, 0, sizeof(unsigned long long), ctx->st));
        ctx->in_rounds = true;
        int32_t r = STF_OK;
        for (int k = 0; k < todo && !r; k++) {
            const int ni = round + k;
            LAUNCH_PDL(ctx, k_round_begin, 1, 32, 0, C0);
            if (ni > 0) { ctx->dev_labels = ctx->inds.as<int32_t>(); ctx->dev_nl = ctx->epoch_ctl.as<int>(); }
            r = spacial_core_fast(ctx, ctx->n, nullptr, 0);
            ctx->dev_labels = nullptr; ctx->dev_nl = nullptr;
            if (!r) r = fast_round_tail(ctx, seed, iteration0 + done + k, true);
            if (!r) LAUNCH_PDL(ctx, k_epoch_e_end, 1, 1024, 0, C0, (int)C_STOP, (int)C_STATUS, (int)C_NHITS, (int)C_ITEMS, (int)C_DIRECT, (int)C_VISITED, (int)C_MOVES, (int)C_NVALID,
                               ctx->rounds_rec.as<unsigned long long>(), k, ni, (const Item16*)ctx->hits2.as<Item16>(), (int64_t)BS_CAP,
                               reinterpret_cast<int*>(ctx->lab_mark.as<uint32_t>()), ctx->n, ctx->inds.as<int32_t>(), ctx->epoch_ctl.as<int>());
        }
        ctx->in_rounds = false;
        if (r) return r;
        CK(cudaMemcpyAsync(rec.data(), ctx->rounds_rec.p, sizeof(unsigned long long) * ROUND_REC * (size_t)(todo + 2), cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctl, ctx->epoch_ctl.p, sizeof(ctl), cudaMemcpyDeviceToHost, ctx->st));
        r = sync_counts(ctx); if (r) return r;
        const int ran = (int)rec[0];
        for (int k = 0; k < ran; k++) if (nhits_per_round) nhits_per_round[done + k] = (int64_t)rec[(size_t)ROUND_REC * (k + 1)];
        done += ran; round += ran;
        const int stop = (int)ctx->h_counts[C_STOP];
        CK(cudaMemsetAsync(C0 + C_STOP, 0, sizeof(unsigned long long), ctx->st));
        ctx->lab_epoch = 0; CK(cudaMemsetAsync(ctx->lab_mark.p, 0, sizeof(uint32_t) * (size_t)ctx->n, ctx->st)); // (the marks were used as first positions)
        if (stop == 2) { // a round needs larger buffers (or the scene left the latency path): grow, go on from that round
            if (!fast_round_verdict(ctx)) FAIL(STF_E_CUDA, "internal: a round was cancelled without a status bit");
            if (ctx->force_big) FAIL(STF_E_STATE, "the scene left the single-scene latency path inside stf_fit_epoch_e");
            if (done < max_rounds) continue;
        }
        if (rounds_done) *rounds_done = done;
        state_out[0] = stop == 2 ? 0 : ctl[3]; state_out[1] = ctl[2]; state_out[2] = ctl[0]; state_out[3] = ctl[1];
        if (round == 0) { state_out[0] = 0; state_out[1] = state_out[2] = state_out[3] = 0; } // not even round 0 ran
        if (inds_out && ctl[0] > 0 && round > 0) {
            if (ctl[0] > inds_cap) FAIL(STF_E_CAPACITY, "label buffer too small");
            CK(cudaMemcpyAsync(inds_out, ctx->inds.p, sizeof(int32_t) * (size_t)ctl[0], cudaMemcpyDeviceToHost, ctx->st));
            CK(cudaStreamSynchronize(ctx->st));
        }
        ctx->ctr.hits = done > 0 && nhits_per_round ? nhits_per_round[done - 1] : 0;
        return STF_OK;
    }
    FAIL(STF_E_CUDA, "internal: stf_fit_epoch_e did not converge");
}

// ---- ground composition for place!/reposition!: deepcopy(mask) with every other tree overlapped
// (qtree_functions.jl:505-518 `for i in 1:length(sortedtrees) if i in indexes continue end; overlap!(ground, sortedtrees[i])`).
// The parallel part of place! runs on the device; the sequential, RNG-driven room search reads the result on the host.
extern "C" int32_t stf_ground_compose(stf_ctx* ctx, int32_t mask_label, int32_t n_excl, const int32_t* excl_labels) {
    NEED_OBJS(); CHECK_LABEL(mask_label);
    if (ctx->multi_scene) FAIL(STF_E_STATE, "ground composition is defined for one scene");
    int mo = mask_label - 1, L = ctx->L;
    std::vector<char> skip((size_t)ctx->n, 0);
    skip[mo] = 1;
    for (int i = 0; i < n_excl; i++) { CHECK_LABEL(excl_labels[i]); skip[excl_labels[i] - 1] = 1; }
    std::vector<int32_t> list; list.reserve(ctx->n);
    for (int o = 0; o < ctx->n; o++) if (!skip[o]) list.push_back(o);
    // the ground = a copy of the mask's pyramid (levels are contiguous per object) with its own metas
    uint32_t base = ctx->h_off[(size_t)mo * L];
    uint64_t total = 0; { int a = ctx->h_om[mo].kh1, b = ctx->h_om[mo].kw1; for (int l = 1; l <= L; l++) { total += (uint64_t)wpc_of(a) * b; a = a / 2 + 1; b = b / 2 + 1; } }
    CK(ctx->gwords.ensure(sizeof(uint32_t) * (size_t)(total + 16))); CK(ctx->glm.ensure(sizeof(LevelMeta) * (size_t)L));
    CK(cudaMemcpyAsync(ctx->gwords.p, ctx->words.as<uint32_t>() + base, sizeof(uint32_t) * (size_t)total, cudaMemcpyDeviceToDevice, ctx->st));
    std::vector<LevelMeta> hm((size_t)L);
    CK(cudaMemcpyAsync(hm.data(), ctx->lm.as<LevelMeta>() + (size_t)mo * L, sizeof(LevelMeta) * (size_t)L, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    for (int l = 0; l < L; l++) hm[l].off -= base;
    CK(stage_h2d(ctx, ctx->glm.p, hm.data(), sizeof(LevelMeta) * (size_t)L));
    if (!list.empty()) {
        CK(ctx->lab4.ensure(sizeof(int32_t) * list.size()));
        CK(stage_h2d(ctx, ctx->lab4.p, list.data(), sizeof(int32_t) * list.size()));
        int blocks = std::min(nblk((int64_t)list.size() * 32, 256), 148 * 8);
        ctx->launches++, k_overlap_level1<<<blocks, 256, 0, ctx->st>>>(ctx->words.as<uint32_t>(), ctx->lm.as<LevelMeta>(), L, ctx->lab4.as<int32_t>(), (int)list.size(),
                                                                  ctx->gwords.as<uint32_t>(), hm[0]);
        KCHECK();
    }
    // buildqtree!(ground): every upper level from the composed level 1 (shifts unchanged: same chain as the mask's)
    int32_t r = build_big_on(ctx, ctx->gwords.as<uint32_t>(), ctx->glm.as<LevelMeta>(), 0, ctx->h_om[mo].kh1, ctx->h_om[mo].kw1, 1); if (r) return r;
    ctx->ground_of = mo;
    CK(cudaStreamSynchronize(ctx->st));
    return STF_OK;
}
extern "C" int32_t stf_ground_level(stf_ctx* ctx, int32_t l, uint8_t* out, int32_t info[5]) {
    NEED_OBJS();
    if (ctx->ground_of < 0) FAIL(STF_E_STATE, "no ground composed");
    if (l < 1 || l > ctx->L) FAIL(STF_E_ARG, "level out of range");
    int mo = ctx->ground_of;
    int kh = ((ctx->h_om[mo].kh1 - 2) >> (l - 1)) + 2, kw = ((ctx->h_om[mo].kw1 - 2) >> (l - 1)) + 2;
    if (l == 1) { kh = ctx->h_om[mo].kh1; kw = ctx->h_om[mo].kw1; }
    LevelMeta m;
    CK(cudaMemcpyAsync(&m, ctx->glm.as<LevelMeta>() + (l - 1), sizeof(m), cudaMemcpyDeviceToHost, ctx->st));
    size_t nb = (size_t)kh * kw;
    if (nb && out) {
        CK(ctx->bytes_tmp.ensure(nb));
        ctx->launches++, k_unpack_level<<<std::min(nblk((int64_t)nb, 256), 148 * 8), 256, 0, ctx->st>>>(ctx->gwords.as<uint32_t>(), ctx->glm.as<LevelMeta>(), ctx->L, 0, l, ctx->bytes_tmp.as<uint8_t>());
        KCHECK();
        CK(cudaMemcpyAsync(out, ctx->bytes_tmp.p, nb, cudaMemcpyDeviceToHost, ctx->st));
    }
    CK(cudaStreamSynchronize(ctx->st));
    if (info) { info[0] = kh; info[1] = kw; info[2] = m.rs; info[3] = m.cs; info[4] = ctx->canvas >> (l - 1); }
    return STF_OK;
}

// ---- item-parallel sharding of one scene over several GPUs.  Every rank holds an identical replica;
// rank r of `nranks` generates and tests the candidate pairs of the sorted records r, r+nranks, ...; the local hit
// lists are exchanged by the caller (one all-gather of fixed-stride buffers per fit iteration, NCCL over NVLink), and
// every rank then runs the same canonical batchsteps on the union, which keeps the replicas bit-identical.
extern "C" int32_t stf_shard_collide(stf_ctx* ctx, int32_t mode, int32_t nl, const int32_t* labels, int32_t rank, int32_t nranks, int64_t* nlocal) {
    NEED_OBJS();
    if (nlocal) *nlocal = 0;
    if (nranks < 1 || rank < 0 || rank >= nranks) FAIL(STF_E_ARG, "bad rank");
    if (!labels) nl = ctx->n;
    ctx->last_result = 0;
    if (ctx->n <= 1 || nl <= 0) { ctx->ctr.hits = 0; return STF_OK; }
    int linked = mode == 1;
    ctx->shard_rank = rank; ctx->shard_n = nranks;
    int32_t r = STF_OK; int64_t nh = 0; bool done = false;
    for (int attempt = 0; attempt < 4 && !done && fast_path_ok(ctx, linked ? ctx->n : nl); attempt++) {
        r = spacial_core_fast(ctx, nl, labels, linked); if (r) break;
        r = sync_counts(ctx); if (r) break;
        int64_t items = (int64_t)ctx->h_counts[C_ITEMS], hits = (int64_t)ctx->h_counts[C_NHITS];
        if (ctx->h_err[1] & STF_ST_RETRY_BIG) {
            cudaMemsetAsync(STATUSP(ctx), 0, sizeof(unsigned long long), ctx->st);
            if (single_cta_sorts(ctx) && !ctx->no_cluster) { ctx->small_off = true; continue; }
            ctx->force_big = true; break;
        }
        if (items > (int64_t)(ctx->items.cap / sizeof(Item16)) && (!ctx->fused_narrow || ctx->keep_items)) { ctx->hint_items = items; continue; }
        if (ctx->fused_narrow && !ctx->keep_items && (int64_t)ctx->h_counts[C_SURV] > (int64_t)(ctx->surv.cap / sizeof(SurvRec))) { ctx->hint_surv = (int64_t)ctx->h_counts[C_SURV]; cudaMemsetAsync(STATUSP(ctx), 0, sizeof(unsigned long long), ctx->st); continue; }
        if (hits > (int64_t)(ctx->hits.cap / sizeof(Item16))) { ctx->force_big = true; break; }
        read_round_counters(ctx); nh = hits; done = true;
    }
    if (!r && !done) r = spacial_core(ctx, nl, labels, linked, &nh);
    ctx->shard_rank = 0; ctx->shard_n = 1;
    if (r) return r;
    ctx->ctr.hits = nh;
    if (nlocal) *nlocal = nh;
    return STF_OK;
}
extern "C" int32_t stf_shard_pack(stf_ctx* ctx, void* send_dev, int64_t stride) {
    NEED_OBJS();
    int64_t n = ctx->ctr.hits;
    if (!send_dev || stride < n + 1) FAIL(STF_E_CAPACITY, "exchange stride too small for the local hit list");
    ctx->launches++, k_shard_pack<<<nblk(std::max<int64_t>(1, n), 256), 256, 0, ctx->st>>>(ctx->hits.as<Item16>(), n, reinterpret_cast<Item16*>(send_dev), stride);
    KCHECK();
    CK(cudaStreamSynchronize(ctx->st));
    return STF_OK;
}
extern "C" int32_t stf_shard_step(stf_ctx* ctx, const void* recv_dev, int32_t nranks, int64_t stride, uint64_t seed, uint64_t iteration, int64_t* nhits) {
    NEED_OBJS();
    if (nhits) *nhits = 0;
    if (!recv_dev || nranks < 1 || stride < 1) FAIL(STF_E_ARG, "bad exchange buffer");
    int64_t cap = (int64_t)nranks * (stride - 1);
    CK(ctx->hits.ensure(sizeof(Item16) * (size_t)std::max<int64_t>(1, cap)));
    ctx->launches++, k_shard_unpack<<<nblk((int64_t)nranks * stride, 256), 256, 0, ctx->st>>>(reinterpret_cast<const Item16*>(recv_dev), nranks, stride, ctx->hits.as<Item16>(), cap,
                                                                                          ctx->counts.as<unsigned long long>() + C_NHITS);
    KCHECK();
    int32_t r = sync_counts(ctx); if (r) return r;
    int64_t nh = (int64_t)ctx->h_counts[C_NHITS], nout = 0;
    if (nh > cap) FAIL(STF_E_ARG, "corrupt exchange headers");
    r = finalize_hits(ctx, nh, 0, false, &nout); if (r) return r; // canonical list order = sorted: identical on every rank
    if (nhits) *nhits = nh;
    ctx->ctr.hits = nh;
    return batchsteps_device(ctx, ctx->hits2.as<Item16>(), nh, seed, iteration, ctx->scene_sorted);
}
// ------------------------------------------------------------------------------------------------ group of GPUs, one process
// SURVEY §8b threading row / §8e-2: ONE host thread drives one context per GPU (the way a Julia host would); the scene is
// replicated, every fit round is split by candidate records, and the local hit lists are exchanged by the GPUs themselves:
// peer stores over NVLink + a device-side flag barrier (k_shard_push / k_shard_gather).  The host only enqueues: no
// collective call, no host synchronisation between the halves of a round — one synchronisation per GPU at the end.
// One HOST THREAD per GPU of the group: enqueueing a round is ~15 launches per device, so a single thread would serialise
// the GPUs on the host side (measured: 2 GPUs slower than 1).  The workers are persistent; a job is a function of the
// device index; after a job a worker spins briefly (the next round usually follows at once) before it sleeps on a
// condition variable.  Thread 0 is the caller's own thread.
struct GroupPool {
    std::vector<std::thread> th;
    std::mutex mu; std::condition_variable cv;
    std::function<void(int)> job;
    std::atomic<unsigned long long> gen{0}; std::atomic<int> pending{0}; std::atomic<bool> quit{false};
    void start(int n) {
        for (int i = 1; i < n; i++) th.emplace_back([this, i] {
            unsigned long long seen = 0;
            for (;;) {
                auto t0 = std::chrono::steady_clock::now();
                while (gen.load(std::memory_order_acquire) == seen && !quit.load(std::memory_order_acquire)) {
                    if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(300)) {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return gen.load(std::memory_order_acquire) != seen || quit.load(std::memory_order_acquire); });
                        break;
                    }
                }
                if (quit.load(std::memory_order_acquire)) return;
                seen = gen.load(std::memory_order_acquire);
                job(i);
                pending.fetch_sub(1, std::memory_order_acq_rel);
            }
        });
    }
    void run(int n, const std::function<void(int)>& f) { // f(i) for every device i, f(0) on the calling thread
        if (n == 1) { f(0); return; }
        job = f;
        pending.store(n - 1, std::memory_order_release);
        { std::lock_guard<std::mutex> lk(mu); gen.fetch_add(1, std::memory_order_acq_rel); }
        cv.notify_all();
        f(0);
        while (pending.load(std::memory_order_acquire) != 0) std::this_thread::yield();
    }
    void stop() {
        { std::lock_guard<std::mutex> lk(mu); quit.store(true, std::memory_order_release); }
        cv.notify_all();
        for (auto& t : th) t.join();
        th.clear();
    }
};
struct stf_group {
    std::vector<stf_ctx*> ctx; std::string err;
    unsigned long long epoch = 0; int64_t seg_cap = 0;
    int64_t nvlink_bytes = 0; // bytes pushed to peer GPUs by the last round (hit records + flags, excluding the local copy)
    bool shared_device = false; // two contexts of the group sit on the same GPU (one-GPU test boxes): see stf_group_fit_round
    GroupPool pool;
};
extern "C" int32_t stf_group_create(int32_t n, const int32_t* devices, stf_group** out) {
    if (!out || n < 1 || n > STF_MAX_GROUP || !devices) return STF_E_ARG;
    *out = nullptr;
    stf_group* g = new stf_group();
    for (int i = 0; i < n; i++) {
        stf_ctx* c = nullptr;
        int32_t r = stf_create(devices[i], &c);
        if (r) { for (stf_ctx* x : g->ctx) stf_destroy(x); delete g; return r; }
        g->ctx.push_back(c);
    }
    for (int i = 0; i < n; i++) // peer access both ways between every pair of distinct devices
        for (int j = 0; j < n; j++) {
            if (devices[i] == devices[j]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
            cudaSetDevice(devices[i]);
            cudaError_t e = can ? cudaDeviceEnablePeerAccess(devices[j], 0) : cudaErrorPeerAccessUnsupported;
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) { for (stf_ctx* x : g->ctx) stf_destroy(x); delete g; return STF_E_CUDA; }
        }
    for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) if (devices[i] == devices[j]) g->shared_device = true;
    g->seg_cap = BS_CAP + 1024;
    for (int i = 0; i < n; i++) {
        stf_ctx* ctx = g->ctx[i];
        cudaSetDevice(ctx->device);
        if (ctx->gather.ensure(sizeof(Item16) * (size_t)n * g->seg_cap) != cudaSuccess || ctx->gflags.ensure(sizeof(unsigned long long) * 2 * STF_MAX_GROUP) != cudaSuccess ||
            ctx->gdone.ensure(sizeof(unsigned int) * 4) != cudaSuccess) { for (stf_ctx* x : g->ctx) stf_destroy(x); delete g; return STF_E_CUDA; }
        cudaMemset(ctx->gflags.p, 0, sizeof(unsigned long long) * 2 * STF_MAX_GROUP); cudaMemset(ctx->gdone.p, 0, sizeof(unsigned int) * 4);
    }
    g->pool.start(n);
    *out = g;
    return STF_OK;
}
extern "C" void stf_group_destroy(stf_group* g) {
    if (!g) return;
    g->pool.stop();
    for (stf_ctx* x : g->ctx) stf_destroy(x);
    delete g;
}
extern "C" int32_t stf_group_size(const stf_group* g) { return g ? (int32_t)g->ctx.size() : 0; }
extern "C" stf_ctx* stf_group_ctx(stf_group* g, int32_t i) { return (g && i >= 0 && i < (int32_t)g->ctx.size()) ? g->ctx[i] : nullptr; }
extern "C" const char* stf_group_last_error(const stf_group* g) { return g ? g->err.c_str() : "no group"; }
extern "C" int64_t stf_group_nvlink_bytes(const stf_group* g) { return g ? g->nvlink_bytes : 0; }
// the same state change on every replica, the GPUs driven in parallel: setshift!/shift! (stf_set_shifts) and reset!(optimiser)
extern "C" int32_t stf_group_set_shifts(stf_group* g, int32_t n, const int32_t* labels, int32_t level, const int32_t* rs, const int32_t* cs, int32_t mode) {
    if (!g) return STF_E_ARG;
    if (mode & (STF_SHIFT_DEVICE_ARRAYS | STF_SHIFT_DEVICE_LABELS)) { g->err = "group calls take host arrays"; return STF_E_ARG; }
    std::vector<int32_t> rc(g->ctx.size(), STF_OK);
    g->pool.run((int)g->ctx.size(), [&](int i) { rc[i] = stf_set_shifts(g->ctx[i], n, labels, level, rs, cs, mode); });
    for (size_t i = 0; i < rc.size(); i++) if (rc[i]) { g->err = g->ctx[i]->err; return rc[i]; }
    return STF_OK;
}
extern "C" int32_t stf_group_reset_velocity(stf_group* g, int32_t n, const int32_t* labels) {
    if (!g) return STF_E_ARG;
    std::vector<int32_t> rc(g->ctx.size(), STF_OK);
    g->pool.run((int)g->ctx.size(), [&](int i) { rc[i] = stf_reset_velocity(g->ctx[i], n, labels); });
    for (size_t i = 0; i < rc.size(); i++) if (rc[i]) { g->err = g->ctx[i]->err; return rc[i]; }
    return STF_OK;
}

#define GFAIL(code, msg) do { g->err = (msg); return (code); } while (0)
extern "C" int32_t stf_group_fit_round(stf_group* g, int32_t mode, int32_t nl, const int32_t* labels, uint64_t seed, uint64_t iteration, int64_t* nhits) {
    if (!g) return STF_E_ARG;
    if (nhits) *nhits = 0;
    const int n = (int)g->ctx.size();
    stf_ctx* c0 = g->ctx[0];
    for (stf_ctx* c : g->ctx) if (c->L == 0 || c->n != c0->n || c->L != c0->L) GFAIL(STF_E_STATE, "every context of the group must hold the same uploaded scene");
    if (!labels) nl = c0->n;
    if (c0->n <= 1 || nl <= 0) return STF_OK;
    const int linked = mode == 1;
    PeerSet P = {}; P.n = n;
    for (int i = 0; i < n; i++) { P.gather[i] = g->ctx[i]->gather.as<Item16>(); P.flags[i] = g->ctx[i]->gflags.as<unsigned long long>(); }
    for (int attempt = 0; attempt < 4; attempt++) {
        if (!fast_path_ok(c0, linked ? c0->n : nl)) GFAIL(STF_E_STATE, "scene too large for the item-parallel latency path of a group (use scene-parallel contexts)");
        const unsigned long long epoch = ++g->epoch;
        std::vector<int32_t> rcs(n, STF_OK);
        // every GPU on its own host thread.  Phase A: broad phase + narrow phase over its share of the records, then the local
        // hits are pushed to all peers.  Phase B: wait for all peers ON THE DEVICE, union of the hit lists, canonical order,
        // batchsteps.  Then the GPU's one synchronisation.
        // Contexts that SHARE a device (a group listed as {0, 0, 0} on a one-GPU box) cannot use the device-side barrier: a
        // spinning gather kernel only ends if the other contexts' kernels run beside it, and kernels of different streams of one
        // device may be serialised (streams aliased onto one hardware queue).  There the halves are separated by a host
        // synchronisation of every context, so the flags are already set when the gather kernels start.
        std::vector<char> pushed_v(n, 0);
        auto fail_of = [&](int i) {
            return [&, i](int32_t code, const char* msg) {
                stf_ctx* ctx = g->ctx[i];
                if (msg) ctx->err = msg;
                rcs[i] = code;
                if (!pushed_v[i]) { ctx->launches++; k_shard_poison<<<1, 32, 0, ctx->st>>>(P, i, epoch); cudaStreamSynchronize(ctx->st); } // the peers must not wait for this GPU forever
            };
        };
        auto phase_a = [&](int i) {
            stf_ctx* ctx = g->ctx[i];
            auto fail = fail_of(i);
            if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(STF_E_CUDA, "cudaSetDevice failed");
            stage_reset(ctx); ctx->after_kernel = false; ctx->last_result = 0;
            ctx->shard_rank = i; ctx->shard_n = n;
            int32_t r = spacial_core_fast(ctx, nl, labels, linked);
            ctx->shard_rank = 0; ctx->shard_n = 1;
            if (r) return fail(r, nullptr);
            FinalizeArgs F = {}; set_overflow_check(ctx, F);
            LAUNCH_PDL(ctx, k_shard_push, 8, 256, 0, (const Item16*)ctx->hits.as<Item16>(), (const unsigned long long*)(ctx->counts.as<unsigned long long>() + C_NHITS), g->seg_cap, P, i, epoch,
                       (const int*)STATUSP(ctx), F.n_items_dev, F.item_cap, ctx->gdone.as<unsigned int>());
            if (cudaGetLastError() != cudaSuccess) return fail(STF_E_CUDA, "k_shard_push launch failed");
            pushed_v[i] = 1;
            if (g->shared_device && cudaStreamSynchronize(ctx->st) != cudaSuccess) return fail(STF_E_CUDA, "synchronisation failed");
        };
        auto phase_b = [&](int i) {
            stf_ctx* ctx = g->ctx[i];
            auto fail = fail_of(i);
            if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(STF_E_CUDA, "cudaSetDevice failed");
            int32_t r = STF_OK;
            if (ctx->hits.ensure(sizeof(Item16) * (size_t)n * g->seg_cap, true, ctx->st) != cudaSuccess) return fail(STF_E_CUDA, "hit buffer allocation failed");
            LAUNCH_PDL(ctx, k_shard_gather, 16, 256, 0, (const Item16*)ctx->gather.as<Item16>(), (const unsigned long long*)ctx->gflags.as<unsigned long long>(), n, g->seg_cap, epoch,
                       ctx->hits.as<Item16>(), (int64_t)(ctx->hits.cap / sizeof(Item16)), ctx->counts.as<unsigned long long>() + C_NHITS, STATUSP(ctx));
            if (cudaGetLastError() != cudaSuccess) return fail(STF_E_CUDA, "k_shard_gather launch failed");
            r = fast_round_tail(ctx, seed, iteration, true); if (r) return fail(r, nullptr);
            r = sync_counts(ctx); if (r) return fail(r, nullptr);
        };
        if (g->shared_device) {
            g->pool.run(n, phase_a);
            g->pool.run(n, [&](int i) { if (pushed_v[i]) phase_b(i); });
        } else g->pool.run(n, [&](int i) { phase_a(i); if (!rcs[i]) phase_b(i); });
        bool again = false;
        for (int i = 0; i < n; i++) if (rcs[i]) GFAIL(rcs[i], g->ctx[i]->err); // (a failed enqueue on one GPU leaves its peers' gather kernels without its flag only if the push never ran: every error above happens after the push or before any launch of the round)
        for (int i = 0; i < n; i++) { cudaSetDevice(g->ctx[i]->device); if (fast_round_verdict(g->ctx[i])) again = true; }
        if (again) { // a list was too small somewhere (every GPU dropped the round: the cancel bit travels with the hits): sizes are grown, run it again
            for (stf_ctx* c : g->ctx) if (c->force_big) GFAIL(STF_E_STATE, "scene too large for the item-parallel latency path of a group (use scene-parallel contexts)");
            continue;
        }
        for (int i = 0; i < n; i++) { stf_ctx* ctx = g->ctx[i]; read_round_counters(ctx); ctx->ctr.moved = (int64_t)ctx->h_counts[C_MOVES]; }
        if (nhits) *nhits = c0->ctr.hits; // the union: identical on every GPU
        g->nvlink_bytes = (int64_t)(n - 1) * ((int64_t)sizeof(Item16) * c0->ctr.hits + 16ll * n); // every hit travels to n-1 peers; 16 B of flags per (source, peer)
        return STF_OK;
    }
    GFAIL(STF_E_CUDA, "internal: the group round did not converge");
}

extern "C" int32_t stf_counters(stf_ctx* ctx, stf_counters_t* out) {
    if (!ctx || !out) return STF_E_ARG;
    *out = ctx->ctr;
    return STF_OK;
}

// ------------------------------------------------------------------------------------------------ per-stage device timers
extern "C" int32_t stf_profile_enable(stf_ctx* ctx, int32_t on) {
    if (!ctx) return STF_E_ARG;
    stage_end(ctx);
    ctx->prof = on != 0;
    return STF_OK;
}
extern "C" int32_t stf_profile_read(stf_ctx* ctx, double* ms, int64_t* calls, int32_t reset) {
    if (!ctx) return STF_E_ARG;
    CK(cudaSetDevice(ctx->device));
    stage_end(ctx);
    CK(cudaStreamSynchronize(ctx->st));
    for (auto& r : ctx->prof_pending) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ctx->prof_ms[r.stage] += t; ctx->prof_calls[r.stage]++; }
        ctx->prof_pool.push_back(r.a); ctx->prof_pool.push_back(r.b);
    }
    ctx->prof_pending.clear();
    for (int i = 0; i < STF_NSTAGES; i++) { if (ms) ms[i] = ctx->prof_ms[i]; if (calls) calls[i] = ctx->prof_calls[i]; }
    if (reset) for (int i = 0; i < STF_NSTAGES; i++) { ctx->prof_ms[i] = 0; ctx->prof_calls[i] = 0; }
    return STF_OK;
}

extern "C" int32_t stf_profile_build_kernels(stf_ctx* ctx, double* ms4) {
    if (!ctx || !ms4) return STF_E_ARG;
    for (int i = 0; i < 4; i++) ms4[i] = ctx->build_ms[i];
    return STF_OK;
}
extern "C" int64_t stf_launch_count(const stf_ctx* ctx) { return ctx ? ctx->launches : 0; }

// debug: timeline of the last k_batchsteps launch (STF_DEBUG_STEP & 16): 4 x u64 nanosecond stamps per step
extern "C" int32_t stf_debug_step_timeline(stf_ctx* ctx, int64_t n, unsigned long long* out) {
    if (!ctx || !ctx->dbg_t.p) return STF_E_STATE;
    CK(cudaMemcpy(out, ctx->dbg_t.p, sizeof(unsigned long long) * 4 * (size_t)n, cudaMemcpyDeviceToHost));
    return STF_OK;
}

#ifdef STF_BS_DEBUG
extern "C" int32_t stf_debug_sort_cycles(long long* out64) { return cudaMemcpyFromSymbol(out64, g_bs_dbg, sizeof(long long) * 64) == cudaSuccess ? 0 : -1; }
#endif
