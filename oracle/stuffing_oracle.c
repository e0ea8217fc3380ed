/* stuffing_oracle.c — CPU ORACLE, TEST INFRASTRUCTURE ONLY (see stuffing_oracle.h).
 * Restates /root/reference/src/{qtrees.jl,qtree_functions.jl,fit.jl,fit_helper.jl};
 * each function cites the lines it follows.  Safe-mode reads, canonical-mode randomness.
 */
#include "stuffing_oracle.h"
#include "../include/stuffing_b200_rng.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ containers */
typedef struct { int32_t l, a, b; } Idx; /* Index = (l,a,b)  qtrees.jl:18 */

typedef struct { Idx* v; int64_t head, n, cap; } IdxQ; /* FIFO used as Vector{Index} */
static void q_init(IdxQ* q) { q->v = NULL; q->head = q->n = q->cap = 0; }
static void q_free(IdxQ* q) { free(q->v); q_init(q); }
static int q_empty(const IdxQ* q) { return q->head == q->n; }
static void q_push(IdxQ* q, Idx i) {
    if (q->n == q->cap) {
        if (q->head > 0) { /* compact */
            memmove(q->v, q->v + q->head, (size_t)(q->n - q->head) * sizeof(Idx));
            q->n -= q->head; q->head = 0;
        }
        if (q->n == q->cap) { q->cap = q->cap ? 2 * q->cap : 64; q->v = (Idx*)realloc(q->v, (size_t)q->cap * sizeof(Idx)); }
    }
    q->v[q->n++] = i;
}
static Idx q_pop(IdxQ* q) { return q->v[q->head++]; }

typedef struct { int32_t* v; int64_t n, cap; } IVec;
static void iv_push(IVec* x, int32_t e) {
    if (x->n == x->cap) { x->cap = x->cap ? 2 * x->cap : 4; x->v = (int32_t*)realloc(x->v, (size_t)x->cap * sizeof(int32_t)); }
    x->v[x->n++] = e;
}
static void iv_pushfirst(IVec* x, int32_t e) {
    iv_push(x, e);
    memmove(x->v + 1, x->v, (size_t)(x->n - 1) * sizeof(int32_t));
    x->v[0] = e;
}

typedef struct { OrcItem* v; int64_t n, cap; } ItemVec;
static void itv_push(ItemVec* x, int i1, int i2, Idx at) {
    if (x->n == x->cap) { x->cap = x->cap ? 2 * x->cap : 256; x->v = (OrcItem*)realloc(x->v, (size_t)x->cap * sizeof(OrcItem)); }
    OrcItem it = { i1, i2, at.l, at.a, at.b };
    x->v[x->n++] = it;
}

/* small sequential generator for the RNG-driven, non-hot-path pieces (place!, random child order) */
typedef struct { uint64_t s; } Rng;
static uint64_t rng_u64(Rng* r) { r->s += 0x9E3779B97F4A7C15ull; return stf_mix64(r->s); }
static double rng_f64(Rng* r) { return (double)(rng_u64(r) >> 11) * (1.0 / 9007199254740992.0); }
static int rng_below(Rng* r, int n) { return (int)(rng_u64(r) % (uint64_t)n); }

/* PERM4  qtrees.jl:12-15 */
static const int8_t PERM4[24][4] = {
    {0,1,2,3},{0,1,3,2},{0,2,1,3},{0,2,3,1},{0,3,1,2},{0,3,2,1},
    {1,0,2,3},{1,0,3,2},{1,2,0,3},{1,2,3,0},{1,3,0,2},{1,3,2,0},
    {2,0,1,3},{2,0,3,1},{2,1,0,3},{2,1,3,0},{2,3,0,1},{2,3,1,0},
    {3,0,1,2},{3,0,2,1},{3,1,0,2},{3,1,2,0},{3,2,0,1},{3,2,1,0}};

/* ------------------------------------------------------------------ index algebra  qtrees.jl:19-49 */
static inline Idx child(Idx i, int n) { Idx c = { i.l - 1, 2 * i.a - (n & 1), 2 * i.b - (n >> 1) }; return c; }
/* `÷` truncates toward zero, like C's `/` */
static inline Idx parent(Idx i) { Idx p = { i.l + 1, (i.a + 1) / 2, (i.b + 1) / 2 }; return p; }
static inline void indexcenter(int l, int a, int b, int64_t* oa, int64_t* ob) {
    if (l == 1) { *oa = a; *ob = b; return; }
    int64_t r = (int64_t)1 << (l - 1), h = (int64_t)1 << (l - 2);
    *oa = r * (a - 1) + h; *ob = r * (b - 1) + h;
}
static inline int childnumber2(Idx anc, Idx des) { /* qtrees.jl:26-29 */
    int64_t o2, o3; indexcenter(anc.l - des.l + 1, anc.a, anc.b, &o2, &o3);
    return ((des.b <= o3) << 1) | (des.a <= o2);
}
static inline int inrange(Idx a, Idx b) { /* qtrees.jl:38-49; caller guarantees a.l >= b.l */
    if (a.l > b.l) {
        int64_t r = (int64_t)1 << (a.l - b.l);
        return (b.a >= r * (a.a - 1) + 1 && b.a <= r * a.a && b.b >= r * (a.b - 1) + 1 && b.b <= r * a.b);
    }
    return a.a == b.a && a.b == b.b;
}

/* ------------------------------------------------------------------ PaddedMat / ShiftedQTree */
typedef struct { int kh, kw, rs, cs, sz; uint8_t* k; } Level; /* payload only: kernel[2:end-2,2:end-2] */
struct OrcTree { int L, sizelevel; uint8_t dflt; Level* lv;
    int built_from, raw_edit; /* bookkeeping of orc_setshifts' skip: level the last rebuild started from; a raw PaddedMat shift edit happened */ };
#define LV(t, l) ((t)->lv[(l) - 1])

static inline int lv_inkernel(const Level* m, int r, int c) { /* inkernelbounds  qtrees.jl:140-149 */
    if (r <= m->rs || c <= m->cs) return 0;
    if (r > m->kh + m->rs || c > m->kw + m->cs) return 0;
    return 1;
}
static inline uint8_t rd(const OrcTree* t, int l, int r, int c) { /* getindex  qtrees.jl:156-162 */
    const Level* m = &LV(t, l);
    if (lv_inkernel(m, r, c)) return m->k[(size_t)(c - m->cs - 1) * m->kh + (r - m->rs - 1)];
    return t->dflt;
}
static inline uint8_t rdi(const OrcTree* t, Idx i) { return rd(t, i.l, i.a, i.b); }

OrcTree* orc_tree_new(const uint8_t* pic, int kh, int kw, int pitch, int canvas, int dflt) {
    /* ShiftedQTree(pic, sz; default) + ShiftedQTree(::PaddedMat)  qtrees.jl:176-200 */
    int L = 1; for (int s = canvas; s > 1; s >>= 1) L++;
    OrcTree* t = (OrcTree*)calloc(1, sizeof(OrcTree));
    t->L = L; t->dflt = (uint8_t)dflt; t->lv = (Level*)calloc((size_t)L, sizeof(Level));
    int m = kh, n = kw, sz = canvas; t->sizelevel = -1; t->built_from = L + 1; /* nothing built yet */
    for (int l = 1; l <= L; l++) {
        Level* v = &LV(t, l);
        v->kh = m; v->kw = n; v->rs = v->cs = 0; v->sz = sz;
        v->k = (uint8_t*)malloc((size_t)m * n + 1);
        memset(v->k, dflt, (size_t)m * n + 1);
        if (l > 1 && t->sizelevel == -1 && m <= 2 && n <= 2) t->sizelevel = l;
        sz /= 2; m = m / 2 + 1; n = n / 2 + 1;
    }
    if (t->sizelevel != -1) t->sizelevel = L; /* qtrees.jl:191-193 */
    for (int c = 0; c < kw; c++) memcpy(LV(t, 1).k + (size_t)c * kh, pic + (size_t)c * pitch, (size_t)kh);
    return t;
}
OrcTree* orc_tree_copy(const OrcTree* s) {
    OrcTree* t = (OrcTree*)calloc(1, sizeof(OrcTree));
    *t = *s; t->lv = (Level*)calloc((size_t)s->L, sizeof(Level));
    for (int l = 1; l <= s->L; l++) {
        LV(t, l) = LV(s, l);
        size_t nb = (size_t)LV(s, l).kh * LV(s, l).kw + 1;
        LV(t, l).k = (uint8_t*)malloc(nb); memcpy(LV(t, l).k, LV(s, l).k, nb);
    }
    return t;
}
void orc_tree_free(OrcTree* t) { if (!t) return; for (int l = 1; l <= t->L; l++) free(LV(t, l).k); free(t->lv); free(t); }
int orc_tree_length(const OrcTree* t) { return t->L; }
int orc_tree_sizelevel(const OrcTree* t) { return t->sizelevel; }
int orc_tree_default(const OrcTree* t) { return t->dflt; }
void orc_tree_level_info(const OrcTree* t, int l, int32_t out[5]) {
    const Level* v = &LV(t, l); out[0] = v->kh; out[1] = v->kw; out[2] = v->rs; out[3] = v->cs; out[4] = v->sz;
}
void orc_tree_level_data(const OrcTree* t, int l, uint8_t* out) { memcpy(out, LV(t, l).k, (size_t)LV(t, l).kh * LV(t, l).kw); }
int orc_tree_get(const OrcTree* t, int l, int a, int b) { return rd(t, l, a, b); }
int orc_inkernelbounds(const OrcTree* t, int l, int a, int b) { return lv_inkernel(&LV(t, l), a, b); }
int orc_tree_set(OrcTree* t, int l, int a, int b, int v) { /* setindex!  qtrees.jl:164-166: payload writes only here */
    Level* m = &LV(t, l);
    if (!lv_inkernel(m, a, b)) return -1;
    m->k[(size_t)(b - m->cs - 1) * m->kh + (a - m->rs - 1)] = (uint8_t)v; return 0;
}
void orc_level_setshift(OrcTree* t, int l, int rs, int cs) { LV(t, l).rs = rs; LV(t, l).cs = cs; t->raw_edit = 1; }

void orc_build(OrcTree* t, int layer) { /* buildqtree!(t::ShiftedQTree, layer)  qtrees.jl:221-237 */
    t->built_from = layer - 1;
    for (int l = layer; l <= t->L; l++) {
        int m2 = LV(t, l - 1).rs / 2, n2 = LV(t, l - 1).cs / 2; /* ÷ toward zero */
        Level* v = &LV(t, l);
        v->rs = m2; v->cs = n2;
        for (int r = 1; r <= v->kh; r++)
            for (int c = 1; c <= v->kw; c++) {
                int R = m2 + r, C = n2 + c; /* qcode  qtrees.jl:50-53 */
                uint8_t q = rd(t, l - 1, 2 * R, 2 * C) | rd(t, l - 1, 2 * R - 1, 2 * C) |
                            rd(t, l - 1, 2 * R, 2 * C - 1) | rd(t, l - 1, 2 * R - 1, 2 * C - 1);
                v->k[(size_t)(c - 1) * v->kh + (r - 1)] = q;
            }
    }
}
void orc_shift(OrcTree* t, int l, int s1, int s2) { /* shift!  qtrees.jl:267-275 */
    for (int i = l; i >= 1; i--) { LV(t, i).rs += s1; LV(t, i).cs += s2; s1 *= 2; s2 *= 2; }
    orc_build(t, l + 1);
}
void orc_setshift(OrcTree* t, int l, int s1, int s2) { /* setshift!  qtrees.jl:277-285 */
    for (int i = l; i >= 1; i--) { LV(t, i).rs = s1; LV(t, i).cs = s2; s1 *= 2; s2 *= 2; }
    orc_build(t, l + 1);
}
/* setshift!(qtrees[labels[i]], level, (rs[i], cs[i])) for a whole list in ONE call (bench scaffolding: the restore of a
 * seeded layout between timed steps; labels == NULL: trees 1..nl).  skip_same != 0: a tree that already sits there is left
 * alone — exact, not an approximation: its last rebuild started at or below `level`, every level <= level has the
 * requested shift and the levels above carry its ÷2 chain, so buildqtree! would rewrite what is there.  The device side
 * applies the same rule (k_build_objects); both arms of bench.py therefore rebuild the same trees.  Returns the number of
 * trees rebuilt. */
int64_t orc_setshifts(OrcTree* const* trees, const int32_t* labels, int nl, int level, const int32_t* rs, const int32_t* cs, int skip_same, int nthreads) {
    int64_t rebuilt = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads) reduction(+ : rebuilt)
    for (int i = 0; i < nl; i++) {
        OrcTree* t = trees[labels ? labels[i] - 1 : i];
        if (skip_same && !t->raw_edit && t->built_from <= level) {
            int same = 1, s1 = rs[i], s2 = cs[i];
            for (int l = level; l >= 1 && same; l--) { same = LV(t, l).rs == s1 && LV(t, l).cs == s2; s1 *= 2; s2 *= 2; }
            for (int l = level + 1; l <= t->L && same; l++) same = LV(t, l).rs == LV(t, l - 1).rs / 2 && LV(t, l).cs == LV(t, l - 1).cs / 2;
            if (same) continue;
        }
        orc_setshift(t, level, rs[i], cs[i]);
        rebuilt++;
    }
    return rebuilt;
}
/* getshift(qtrees[labels[i]], level) for a list  qtrees.jl:288 */
void orc_getshifts(OrcTree* const* trees, const int32_t* labels, int nl, int level, int32_t* rs, int32_t* cs) {
    for (int i = 0; i < nl; i++) { const OrcTree* t = trees[labels ? labels[i] - 1 : i]; rs[i] = LV(t, level).rs; cs[i] = LV(t, level).cs; }
}
void orc_shift_axis(OrcTree* t, int l, int st, int axis, int set) { /* qtrees.jl:238-265 */
    for (int i = l; i >= 1; i--) {
        int* p = axis == 0 ? &LV(t, i).rs : &LV(t, i).cs;
        *p = set ? st : *p + st; st *= 2;
    }
    orc_build(t, l + 1);
}
int orc_testqtree(const OrcTree* t) { /* utils.jl:7-24 over the canvas of each level */
    for (int l = 2; l <= t->L; l++) {
        int s = LV(t, l).sz;
        for (int i = 1; i <= s; i++) for (int j = 1; j <= s; j++) {
            uint8_t c[4] = { rd(t, l - 1, 2 * i, 2 * j), rd(t, l - 1, 2 * i - 1, 2 * j), rd(t, l - 1, 2 * i, 2 * j - 1), rd(t, l - 1, 2 * i - 1, 2 * j - 1) };
            int allf = 1, alle = 1; for (int k = 0; k < 4; k++) { allf &= c[k] == ORC_FULL; alle &= c[k] == ORC_EMPTY; }
            uint8_t q = rd(t, l, i, j);
            if (q == ORC_FULL) { if (!allf) return l; }
            else if (q == ORC_EMPTY) { if (!alle) return l; }
            else if (q == ORC_MIX) { if (allf || alle) return l; }
            else return -l;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ pair collision */
static Idx dfs(const OrcTree* q1, const OrcTree* q2, Idx i) { /* collision_dfs  qtree_functions.jl:2-20 */
    uint8_t code = rdi(q1, i) & rdi(q2, i);
    if (code == ORC_FULL) return i;
    Idx r = { -i.l, i.a, i.b };
    if (code == ORC_MIX && i.l > 1)
        for (int cn = 0; cn < 4; cn++) { r = dfs(q1, q2, child(i, cn)); if (r.l > 0) return r; }
    return r;
}
void orc_collision_dfs(const OrcTree* q1, const OrcTree* q2, int l, int a, int b, int32_t out[3]) {
    Idx i = { l, a, b }; Idx r = dfs(q1, q2, i); out[0] = r.l; out[1] = r.a; out[2] = r.b;
}
/* _collision_randbfs  qtree_functions.jl:21-42 */
static Idx randbfs(const OrcTree* q1, const OrcTree* q2, IdxQ* q, Rng* rng, int64_t* visited) {
    Idx i = q->v[q->head];
    int64_t nv = 0;
    while (!q_empty(q)) {
        i = q_pop(q);
        const int8_t* perm = rng ? PERM4[rng_below(rng, 24)] : PERM4[0];
        for (int k = 0; k < 4; k++) {
            Idx ci = child(i, perm[k]);
            uint8_t code = rdi(q1, ci) & rdi(q2, ci); nv++;
            if (code == ORC_FULL) { if (visited) *visited += nv; return ci; }
            if (code == ORC_MIX && ci.l > 1) q_push(q, ci);
        }
    }
    if (visited) *visited += nv;
    Idx r = { -i.l, i.a, i.b }; return r;
}
void orc_collision_bfs(const OrcTree* q1, const OrcTree* q2, int l, int a, int b, uint64_t perm_seed, int32_t out[3], int64_t* visited) {
    IdxQ q; q_init(&q); Idx at = { l, a, b }; q_push(&q, at);
    Rng rng = { perm_seed };
    Idx r = randbfs(q1, q2, &q, perm_seed ? &rng : NULL, visited);
    out[0] = r.l; out[1] = r.a; out[2] = r.b; q_free(&q);
}
void orc_collision(const OrcTree* q1, const OrcTree* q2, uint64_t perm_seed, int32_t out[3]) { /* qtree_functions.jl:43-51 */
    int l = q1->L;
    if (lv_inkernel(&LV(q1, l), 1, 1) && lv_inkernel(&LV(q2, l), 1, 1)) orc_collision_bfs(q1, q2, l, 1, 1, perm_seed, out, NULL);
    else { out[0] = -l; out[1] = 1; out[2] = 1; }
}

/* ------------------------------------------------------------------ batch narrow phase */
/* _totalcollisions_native(qtrees, coitems, colist)  qtree_functions.jl:89-110.  The reference
 * spawns one task per chunk and pushes hits under a SpinLock in arrival order; here hits keep
 * item order so that the result is deterministic. */
int64_t orc_collide_items(OrcTree* const* trees, const OrcItem* items, int64_t n, uint64_t perm_seed,
                          int nthreads, OrcItem* out, int64_t cap, int64_t* visited) {
    OrcItem* res = (OrcItem*)malloc((size_t)(n > 0 ? n : 1) * sizeof(OrcItem));
    int64_t vis = 0;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1) reduction(+ : vis)
#endif
    {
        IdxQ q; q_init(&q);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t k = 0; k < n; k++) {
            q.head = q.n = 0;
            Idx at = { items[k].l, items[k].a, items[k].b }; q_push(&q, at);
            Rng rng = { perm_seed + (uint64_t)k * 0x632BE59BD9B4E019ull };
            Idx r = randbfs(trees[items[k].i1 - 1], trees[items[k].i2 - 1], &q, perm_seed ? &rng : NULL, &vis);
            OrcItem o = { items[k].i1, items[k].i2, r.l, r.a, r.b }; res[k] = o;
        }
        q_free(&q);
    }
    int64_t m = 0;
    for (int64_t k = 0; k < n; k++) if (res[k].l >= 0) { if (m < cap) out[m] = res[k]; m++; }
    free(res);
    if (visited) *visited += vis;
    return m <= cap ? m : -m;
}

int64_t orc_totalcollisions_native(OrcTree* const* trees, int n, const int32_t* labels, int nl,
                                   uint64_t perm_seed, OrcItem* out, int64_t cap) {
    /* totalcollisions_native + the label-list method  qtree_functions.jl:111-131 */
    if (n <= 1) return 0;
    int L = trees[0]->L;
    IVec lb = { 0 };
    for (int i = 0; i < (labels ? nl : n); i++) {
        int x = labels ? labels[i] : i + 1;
        if (lv_inkernel(&LV(trees[x - 1], L), 1, 1)) iv_push(&lb, x);
    }
    ItemVec items = { 0 }; Idx at = { L, 1, 1 };
    for (int64_t i = 0; i < lb.n; i++) for (int64_t j = lb.n - 1; j > i; j--) itv_push(&items, lb.v[i], lb.v[j], at);
    int64_t r = orc_collide_items(trees, items.v, items.n, perm_seed, 1, out, cap, NULL);
    free(items.v); free(lb.v);
    return r;
}

/* ------------------------------------------------------------------ locate!  qtree_functions.jl:141-189 */
static Idx positionlower(const OrcTree* qt, Idx ind) {
    while (ind.l > 2) {
        Idx cind = ind; int flag = 0;
        for (int i = 0; i < 4; i++) {
            Idx ci = child(ind, i);
            if (rdi(qt, ci) != ORC_EMPTY) { if (flag) return ind; flag = 1; cind = ci; }
        }
        if (!flag) return ind; /* non-EMPTY node without a non-EMPTY child (inconsistent pyramid): the
                                  reference would spin forever on `ind = cind`; oracle and device stop here */
        ind = cind;
    }
    return ind;
}
static int locate_cells(const OrcTree* qt, Idx inds[4]) {
    int l = qt->sizelevel; /* == length(qt), see qtrees.jl:191-193 */
    int rs = LV(qt, l).rs, cs = LV(qt, l).cs, n = 0;
    static const int dr[4] = { 1, 1, 2, 2 }, dc[4] = { 1, 2, 1, 2 };
    for (int k = 0; k < 4; k++)
        if (rd(qt, l, rs + dr[k], cs + dc[k]) != ORC_EMPTY) { Idx i = { l, rs + dr[k], cs + dc[k] }; inds[n++] = i; }
    if (n > 2 && l < qt->L) { /* dead branch (sizelevel == L); kept for fidelity  :171-183 */
        l = l + 1; rs = LV(qt, l).rs; cs = LV(qt, l).cs;
        Idx i2[4]; int n2 = 0;
        for (int k = 0; k < 4; k++)
            if (rd(qt, l, rs + dr[k], cs + dc[k]) != ORC_EMPTY) { Idx i = { l, rs + dr[k], cs + dc[k] }; i2[n2++] = i; }
        if (n2 < n - 1) { n = n2; memcpy(inds, i2, sizeof(i2)); }
    }
    for (int k = 0; k < n; k++) inds[k] = positionlower(qt, inds[k]);
    return n;
}

/* HashSpacialQTree  qtrees.jl:411-422 as an insertion-ordered open-addressing map */
typedef struct { Idx key; IVec labels; } Cell;
typedef struct { Cell* cells; int64_t n, cap; int64_t* slots; int64_t nslots; } HashTree;
static uint64_t idx_hash(Idx k) { return stf_mix64(((uint64_t)(uint32_t)k.l << 58) ^ ((uint64_t)(uint32_t)k.a << 29) ^ (uint64_t)(uint32_t)k.b); }
static void ht_init(HashTree* h) { memset(h, 0, sizeof(*h)); h->nslots = 1024; h->slots = (int64_t*)malloc(sizeof(int64_t) * (size_t)h->nslots); memset(h->slots, 0xff, sizeof(int64_t) * (size_t)h->nslots); }
static void ht_free(HashTree* h) { for (int64_t i = 0; i < h->n; i++) free(h->cells[i].labels.v); free(h->cells); free(h->slots); }
static int64_t ht_find(const HashTree* h, Idx k) {
    uint64_t s = idx_hash(k) & (uint64_t)(h->nslots - 1);
    while (h->slots[s] >= 0) {
        const Cell* c = &h->cells[h->slots[s]];
        if (c->key.l == k.l && c->key.a == k.a && c->key.b == k.b) return h->slots[s];
        s = (s + 1) & (uint64_t)(h->nslots - 1);
    }
    return -1;
}
static void ht_grow(HashTree* h) {
    h->nslots *= 2; h->slots = (int64_t*)realloc(h->slots, sizeof(int64_t) * (size_t)h->nslots);
    memset(h->slots, 0xff, sizeof(int64_t) * (size_t)h->nslots);
    for (int64_t i = 0; i < h->n; i++) {
        uint64_t s = idx_hash(h->cells[i].key) & (uint64_t)(h->nslots - 1);
        while (h->slots[s] >= 0) s = (s + 1) & (uint64_t)(h->nslots - 1);
        h->slots[s] = i;
    }
}
static void ht_push(HashTree* h, Idx k, int label) { /* push!(get!(Vector{Int}, t.qtree, ind), label) */
    int64_t c = ht_find(h, k);
    if (c < 0) {
        if (h->n == h->cap) { h->cap = h->cap ? 2 * h->cap : 256; h->cells = (Cell*)realloc(h->cells, (size_t)h->cap * sizeof(Cell)); }
        c = h->n++; h->cells[c].key = k; memset(&h->cells[c].labels, 0, sizeof(IVec));
        if (2 * h->n > h->nslots) ht_grow(h);
        else { uint64_t s = idx_hash(k) & (uint64_t)(h->nslots - 1); while (h->slots[s] >= 0) s = (s + 1) & (uint64_t)(h->nslots - 1); h->slots[s] = c; }
    }
    iv_push(&h->cells[c].labels, label);
}
static void ht_locate(HashTree* h, OrcTree* const* trees, int n, const int32_t* labels, int nl) { /* locate!  :190-201 */
    for (int i = 0; i < (labels ? nl : n); i++) {
        int x = labels ? labels[i] : i + 1;
        Idx inds[4]; int m = locate_cells(trees[x - 1], inds);
        for (int k = 0; k < m; k++) ht_push(h, inds[k], x);
    }
}
int64_t orc_locate_hash(OrcTree* const* trees, int n, const int32_t* labels, int nl, OrcCellRec* out, int64_t cap) {
    HashTree h; ht_init(&h); ht_locate(&h, trees, n, labels, nl);
    int64_t m = 0;
    for (int64_t c = 0; c < h.n; c++) for (int64_t j = 0; j < h.cells[c].labels.n; j++) {
        if (m < cap) { OrcCellRec r = { h.cells[c].key.l, h.cells[c].key.a, h.cells[c].key.b, h.cells[c].labels.v[j] }; out[m] = r; }
        m++;
    }
    ht_free(&h);
    return m <= cap ? m : -m;
}

/* collisions_boundsfilter  qtree_functions.jl:203-223 */
static void boundsfilter1(OrcTree* const* trees, Idx sp, const int32_t* low, int64_t nlow, int hlb, ItemVec* itemlist, ItemVec* colist) {
    if (lv_inkernel(&LV(trees[hlb - 1], sp.l), sp.a, sp.b)) {
        for (int64_t k = 0; k < nlow; k++) itv_push(itemlist, low[k], hlb, sp);
    } else if (trees[hlb - 1]->dflt == ORC_FULL) {
        for (int64_t k = 0; k < nlow; k++) if (rdi(trees[low[k] - 1], sp) != ORC_EMPTY) itv_push(colist, low[k], hlb, sp);
    }
}

static int item_less(const OrcItem* x, const OrcItem* y) {
    if (x->i1 != y->i1) return x->i1 < y->i1; if (x->i2 != y->i2) return x->i2 < y->i2;
    if (x->l != y->l) return x->l < y->l; if (x->a != y->a) return x->a < y->a; return x->b < y->b;
}
static int item_cmp(const void* x, const void* y) {
    const OrcItem* p = (const OrcItem*)x; const OrcItem* q = (const OrcItem*)y;
    return item_less(p, q) ? -1 : (item_less(q, p) ? 1 : 0);
}
/* unique!(i->minmax(first(i)...), sort!(r))  qtree_functions.jl:252: sort, then keep the
 * first occurrence of every unordered pair (Julia's unique!(f, A) is global, not adjacent-only) */
void orc_sort_unique(OrcItem* r, int64_t* n, int unique) {
    qsort(r, (size_t)*n, sizeof(OrcItem), item_cmp);
    if (!unique) return;
    HashTree seen; ht_init(&seen);
    int64_t m = 0;
    for (int64_t k = 0; k < *n; k++) {
        Idx key = { 0, r[k].i1 < r[k].i2 ? r[k].i1 : r[k].i2, r[k].i1 < r[k].i2 ? r[k].i2 : r[k].i1 };
        if (ht_find(&seen, key) >= 0) continue;
        ht_push(&seen, key, 0); r[m++] = r[k];
    }
    ht_free(&seen); *n = m;
}

static int64_t finish(OrcTree* const* trees, ItemVec* itemlist, ItemVec* colist, int unique, uint64_t perm_seed, int nthreads,
                      OrcItem* out, int64_t cap, OrcItem* items_out, int64_t items_cap, int64_t* n_items, int64_t* n_direct, int64_t* visited) {
    if (n_items) *n_items = itemlist->n;
    if (n_direct) *n_direct = colist->n;
    if (items_out) for (int64_t k = 0; k < itemlist->n && k < items_cap; k++) items_out[k] = itemlist->v[k];
    /* r = _totalcollisions_native(qtrees, itemlist, colist): BFS hits are appended after the direct hits */
    int64_t total_cap = colist->n + itemlist->n + 1;
    OrcItem* r = (OrcItem*)malloc((size_t)total_cap * sizeof(OrcItem));
    memcpy(r, colist->v, (size_t)colist->n * sizeof(OrcItem));
    int64_t nb = orc_collide_items(trees, itemlist->v, itemlist->n, perm_seed, nthreads, r + colist->n, itemlist->n, visited);
    int64_t m = colist->n + nb;
    if (unique) orc_sort_unique(r, &m, 1);
    for (int64_t k = 0; k < m && k < cap; k++) out[k] = r[k];
    free(r); free(itemlist->v); free(colist->v);
    return m <= cap ? m : -m;
}

int64_t orc_totalcollisions_spacial(OrcTree* const* trees, int n, const int32_t* labels, int nl, int unique, uint64_t perm_seed, int nthreads,
                                    OrcItem* out, int64_t cap, OrcItem* items_out, int64_t items_cap, int64_t* n_items, int64_t* n_direct, int64_t* visited) {
    /* totalcollisions_spacial  qtree_functions.jl:225-263 */
    if (n_items) *n_items = 0;
    if (n_direct) *n_direct = 0;
    if (n <= 1) return 0;
    int nlevel = trees[0]->L;
    HashTree h; ht_init(&h); ht_locate(&h, trees, n, labels, nl);
    ItemVec itemlist = { 0 }, colist = { 0 };
    for (int64_t c = 0; c < h.n; c++) { /* for spindex in keys(spqtree): insertion order here */
        Idx sp = h.cells[c].key; const IVec* lbs = &h.cells[c].labels;
        if (lbs->n > 1)
            for (int64_t i = 0; i < lbs->n; i++) for (int64_t j = lbs->n - 1; j > i; j--) itv_push(&itemlist, lbs->v[i], lbs->v[j], sp);
        Idx p = sp;
        for (;;) {
            p = parent(p);
            if (p.l > nlevel) break;
            int64_t pc = ht_find(&h, p);
            if (pc >= 0) for (int64_t k = 0; k < h.cells[pc].labels.n; k++)
                boundsfilter1(trees, sp, lbs->v, lbs->n, h.cells[pc].labels.v[k], &itemlist, &colist);
        }
    }
    ht_free(&h);
    return finish(trees, &itemlist, &colist, unique, perm_seed, nthreads, out, cap, items_out, items_cap, n_items, n_direct, visited);
}

/* ------------------------------------------------------------------ LinkedSpacialQTree  qtrees.jl:424-552 */
typedef struct LNode { Idx index; struct LNode* parent; struct LNode* ch[4]; IVec list; /* front = most recent */ } LNode;
struct OrcLinked { LNode* root; int has_root; LNode*** label_nodes; int* label_n; int nlabels; };
static LNode* ln_new(LNode* parent, Idx index) { LNode* n = (LNode*)calloc(1, sizeof(LNode)); n->parent = parent; n->index = index; return n; }
static void ln_free(LNode* n) { if (!n) return; for (int i = 0; i < 4; i++) ln_free(n->ch[i]); free(n->list.v); free(n); }
OrcLinked* orc_linked_new(int L) {
    OrcLinked* t = (OrcLinked*)calloc(1, sizeof(OrcLinked));
    Idx r = { L, 1, 1 }; t->root = ln_new(NULL, r); t->has_root = L > 0; return t;
}
void orc_linked_free(OrcLinked* t) {
    if (!t) return; ln_free(t->root);
    for (int i = 0; i < t->nlabels; i++) free(t->label_nodes[i]);
    free(t->label_nodes); free(t->label_n); free(t);
}
static void lk_reserve(OrcLinked* t, int label) {
    if (label <= t->nlabels) return;
    int nn = label > 2 * t->nlabels ? label : 2 * t->nlabels;
    t->label_nodes = (LNode***)realloc(t->label_nodes, sizeof(LNode**) * (size_t)nn);
    t->label_n = (int*)realloc(t->label_n, sizeof(int) * (size_t)nn);
    for (int i = t->nlabels; i < nn; i++) { t->label_nodes[i] = (LNode**)calloc(4, sizeof(LNode*)); t->label_n[i] = 0; }
    t->nlabels = nn;
}
static void lk_list_remove(IVec* l, int label) {
    for (int64_t k = 0; k < l->n; k++) if (l->v[k] == label) { memmove(l->v + k, l->v + k + 1, (size_t)(l->n - k - 1) * sizeof(int32_t)); l->n--; return; }
}
void orc_linked_clear(OrcLinked* t, int label) { /* clear!  :530-538 */
    if (label > t->nlabels) return;
    for (int k = 0; k < t->label_n[label - 1]; k++) lk_list_remove(&t->label_nodes[label - 1][k]->list, label);
    t->label_n[label - 1] = 0;
}
static void lk_push(OrcLinked* t, Idx ind, int label) { /* push!  :501-524 */
    if (!t->has_root) return;
    LNode* tn = t->root; Idx aind = tn->index;
    if (aind.l < ind.l || !inrange(aind, ind)) return;
    for (;;) {
        aind = tn->index;
        if (aind.l <= ind.l) break;
        int cn = childnumber2(aind, ind);
        if (!tn->ch[cn]) tn->ch[cn] = ln_new(tn, child(aind, cn));
        tn = tn->ch[cn];
    }
    iv_pushfirst(&tn->list, label);
    lk_reserve(t, label);
    t->label_nodes[label - 1][t->label_n[label - 1]++] = tn;
}
void orc_linked_locate(OrcLinked* t, OrcTree* const* trees, const int32_t* labels, int nl) { /* locate! + update!  :159-201, qtrees.jl:405-410 */
    for (int i = 0; i < nl; i++) {
        int x = labels[i];
        Idx inds[4]; int m = locate_cells(trees[x - 1], inds);
        lk_reserve(t, x);
        orc_linked_clear(t, x);
        for (int k = 0; k < m; k++) lk_push(t, inds[k], x);
    }
}
static void lk_collect(const LNode* n, OrcCellRec* out, int64_t cap, int64_t* m) { /* collect_tree  :542-552 */
    if (!n) return;
    for (int64_t k = 0; k < n->list.n; k++) { if (*m < cap) { OrcCellRec r = { n->index.l, n->index.a, n->index.b, n->list.v[k] }; out[*m] = r; } (*m)++; }
    for (int i = 0; i < 4; i++) lk_collect(n->ch[i], out, cap, m);
}
int64_t orc_linked_collect(const OrcLinked* t, OrcCellRec* out, int64_t cap) { int64_t m = 0; lk_collect(t->root, out, cap, &m); return m <= cap ? m : -m; }

static void lk_descend(OrcTree* const* trees, const LNode* tn, int label, ItemVec* itemlist, ItemVec* colist) {
    /* the stack walk of :302-324 (pruning of empty tree nodes does not change results and is omitted) */
    for (int i = 0; i < 4; i++) {
        const LNode* c = tn->ch[i];
        if (!c) continue;
        if (c->list.n > 0) boundsfilter1(trees, c->index, c->list.v, c->list.n, label, itemlist, colist);
        lk_descend(trees, c, label, itemlist, colist);
    }
}
int64_t orc_partialcollisions(OrcTree* const* trees, int n, OrcLinked* t, const int32_t* labels, int nl, int unique, uint64_t perm_seed, int nthreads,
                              OrcItem* out, int64_t cap, OrcItem* items_out, int64_t items_cap, int64_t* n_items, int64_t* n_direct) {
    /* partialcollisions  qtree_functions.jl:277-330 */
    ItemVec itemlist = { 0 }, colist = { 0 };
    uint8_t* inlabels = (uint8_t*)calloc((size_t)n + 2, 1);
    for (int i = 0; i < nl; i++) inlabels[labels[i]] = 1;
    orc_linked_locate(t, trees, labels, nl);
    for (int i = 0; i < nl; i++) {
        int label = labels[i];
        if (label > t->nlabels) continue;
        for (int k = 0; k < t->label_n[label - 1]; k++) {
            const LNode* treenode = t->label_nodes[label - 1][k];
            Idx sp = treenode->index;
            int64_t pos = 0; while (treenode->list.v[pos] != label) pos++;
            for (int64_t j = pos + 1; j < treenode->list.n; j++) itv_push(&itemlist, label, treenode->list.v[j], sp); /* :293 */
            for (const LNode* tn = treenode->parent; tn; tn = tn->parent) /* :295-301 */
                for (int64_t j = 0; j < tn->list.n; j++) {
                    int plb = tn->list.v[j];
                    if (!inlabels[plb]) { int32_t lo = label; boundsfilter1(trees, sp, &lo, 1, plb, &itemlist, &colist); }
                }
            lk_descend(trees, treenode, label, &itemlist, &colist);
        }
    }
    free(inlabels);
    return finish(trees, &itemlist, &colist, unique, perm_seed, nthreads, out, cap, items_out, items_cap, n_items, n_direct, NULL);
}

/* ------------------------------------------------------------------ fit step */
static inline int decode2(uint8_t c) { static const int T[4] = { 0, 0, 2, 1 }; return T[c & 3]; } /* fit_helper.jl:42 */
void orc_grad2d(const OrcTree* t, int l, int a, int b, int32_t out[2]) { /* fit_helper.jl:51-66 */
#define D(r, c) decode2(rd(t, l, (r), (c)))
    int diag = -D(a - 1, b - 1) + D(a + 1, b + 1);
    int cdiag = -D(a - 1, b + 1) + D(a + 1, b - 1);
    out[0] = diag + cdiag - D(a - 1, b) + D(a + 1, b);
    out[1] = diag - cdiag - D(a, b - 1) + D(a, b + 1);
#undef D
}
int orc_intlog2(double x) { int64_t b; memcpy(&b, &x, 8); return (int)((b >> 52) - 1023); } /* fit_helper.jl:74-78 */

struct OrcOpt { int kind; double eta, rho; int n; double* v; uint8_t* has; };
OrcOpt* orc_opt_new(int kind, double eta, double rho, int n) {
    OrcOpt* o = (OrcOpt*)calloc(1, sizeof(OrcOpt)); o->kind = kind; o->eta = eta; o->rho = rho; o->n = n;
    o->v = (double*)calloc((size_t)2 * n + 2, sizeof(double)); o->has = (uint8_t*)calloc((size_t)n + 1, 1); return o;
}
void orc_opt_free(OrcOpt* o) { if (!o) return; free(o->v); free(o->has); free(o); }
void orc_opt_set_eta(OrcOpt* o, double eta) { o->eta = eta; }
void orc_opt_reset(OrcOpt* o, int label) { o->has[label - 1] = 0; }
void orc_opt_velocity(const OrcOpt* o, int label, double out[2], int32_t* has) { out[0] = o->v[2 * (label - 1)]; out[1] = o->v[2 * (label - 1) + 1]; *has = o->has[label - 1]; }
static void opt_apply(OrcOpt* o, int label, double d[2]) {
    if (!o || o->kind == 0) { d[0] = d[0] / 6; d[1] = d[1] / 6; return; }          /* fit.jl:25 default */
    if (o->kind == 1) { d[0] = o->eta * d[0]; d[1] = o->eta * d[1]; return; }       /* SGD  fit_helper.jl:35 */
    double* v = o->v + 2 * (label - 1);                                              /* Momentum  :19-26 */
    if (!o->has[label - 1]) { v[0] = d[0]; v[1] = d[1]; o->has[label - 1] = 1; }     /* v = get!(velocity, x, Δ) */
    for (int k = 0; k < 2; k++) {
        volatile double t1 = o->rho * v[k], t2 = (1 - o->rho) * d[k];               /* no FMA contraction */
        v[k] = t1 + t2; d[k] = o->eta * v[k];
    }
}
static const double DIAG[4][2] = { { 1., -1. }, { -1., 1. }, { -1., -1. }, { 1., 1. } };
static void move_(OrcTree* qt, double d[2], uint64_t seed, uint64_t iter, uint64_t k, uint32_t slot0) { /* move!  fit.jl:11-23 */
    if (stf_rand_f64(seed, iter, k, slot0) < 0.1) { int j = stf_rand_diag(seed, iter, k, slot0 + 1); d[0] += DIAG[j][0]; d[1] += DIAG[j][1]; }
    if (-1 < d[0] && d[0] < 1 && -1 < d[1] && d[1] < 1) { int j = stf_rand_diag(seed, iter, k, slot0 + 2); d[0] = DIAG[j][0]; d[1] = DIAG[j][1]; }
    double dm = fabs(d[0]) > fabs(d[1]) ? fabs(d[0]) : fabs(d[1]);
    int u = orc_intlog2(dm);
    int64_t p = (int64_t)1 << u;
    int s1 = (int)((int64_t)trunc(d[0]) / p), s2 = (int)((int64_t)trunc(d[1]) / p);
    orc_shift(qt, qt->L < 1 + u ? qt->L : 1 + u, s1, s2);
}
static void step_(OrcTree* const* trees, int i1, int i2, Idx cp, OrcOpt* opt, uint64_t seed, uint64_t iter, uint64_t k) { /* step!  fit.jl:25-48 */
    OrcTree *t1 = trees[i1 - 1], *t2 = trees[i2 - 1];
    double s1 = (double)(LV(t1, 1).kh + LV(t1, 1).kw), s2 = (double)(LV(t2, 1).kh + LV(t2, 1).kw);
    double ks1 = s1 * s1 * s1, ks2 = s2 * s2 * s2; /* exact: integers < 2^53 */
    double ll = (double)((int64_t)1 << (cp.l - 1));
    int32_t g1[2], g2[2]; orc_grad2d(t1, cp.l, cp.a, cp.b, g1); orc_grad2d(t2, cp.l, cp.a, cp.b, g2);
    double d1[2] = { ll * g1[0], ll * g1[1] }, d2[2] = { ll * g2[0], ll * g2[1] };
    int move1 = stf_rand_f64(seed, iter, k, 0) < ks2 / ks1;
    int move2 = stf_rand_f64(seed, iter, k, 1) < ks1 / ks2;
    uint32_t slot = 2;
    if (move1) {
        double d[2] = { d1[0], d1[1] };
        if (!move2) { d[0] -= d2[0]; d[1] -= d2[1]; }
        opt_apply(opt, i1, d); move_(t1, d, seed, iter, k, slot); slot += 3;
    }
    if (move2) {
        double d[2] = { d2[0], d2[1] };
        if (!move1) { d[0] -= d1[0]; d[1] -= d1[1]; }
        opt_apply(opt, i2, d); move_(t2, d, seed, iter, k, slot);
    }
}
static void step_mask_(OrcTree* const* trees, int im, int i2, Idx cp, OrcOpt* opt, uint64_t seed, uint64_t iter, uint64_t k) { /* fit.jl:49-56 */
    OrcTree *mask = trees[im - 1], *t2 = trees[i2 - 1];
    double ll = (double)((int64_t)1 << (cp.l - 1));
    int32_t g1[2], g2[2]; orc_grad2d(mask, cp.l, cp.a, cp.b, g1); orc_grad2d(t2, cp.l, cp.a, cp.b, g2);
    double d[2] = { (ll * g2[0] - ll * g1[0]) / 2, (ll * g2[1] - ll * g1[1]) / 2 };
    opt_apply(opt, i2, d); move_(t2, d, seed, iter, k, 2);
}
void orc_batchsteps_masks(OrcTree* const* trees, int n, const OrcItem* colist, int64_t nh, OrcOpt* opt, uint64_t seed, uint64_t iter,
                          const uint8_t* is_mask) {
    /* batchsteps! + step_index!  fit.jl:58-75.  The reference shuffles colist first; canonical mode
     * takes the list order as the (already arbitrary) order.  is_mask == NULL: label 1 is the mask
     * (fit.jl:60-63); a non-NULL array generalises that convention to batches of scenes. */
    (void)n;
    for (int64_t k = 0; k < nh; k++) {
        Idx cp = { colist[k].l, colist[k].a, colist[k].b };
        int i1 = colist[k].i1, i2 = colist[k].i2;
        int m1 = is_mask ? is_mask[i1 - 1] : i1 == 1, m2 = is_mask ? is_mask[i2 - 1] : i2 == 1;
        if (m1) step_mask_(trees, i1, i2, cp, opt, seed, iter, (uint64_t)k);
        else if (m2) step_mask_(trees, i2, i1, cp, opt, seed, iter, (uint64_t)k);
        else step_(trees, i1, i2, cp, opt, seed, iter, (uint64_t)k);
    }
}
void orc_batchsteps(OrcTree* const* trees, int n, const OrcItem* colist, int64_t nh, OrcOpt* opt, uint64_t seed, uint64_t iter) {
    orc_batchsteps_masks(trees, n, colist, nh, opt, seed, iter, NULL);
}

/* ------------------------------------------------------------------ place! (input generation) */
static inline void wr_raw(OrcTree* t, Idx i, uint8_t v) { orc_tree_set(t, i.l, i.a, i.b, v); }
static void overlap_(OrcTree* t1, const OrcTree* t2, Idx ind) { /* _overlap!  qtree_functions.jl:476-488 */
    if (!(rdi(t1, ind) == ORC_FULL || rdi(t2, ind) == ORC_EMPTY)) {
        if (ind.l == 1) wr_raw(t1, ind, rdi(t2, ind));
        else {
            for (int ci = 0; ci < 4; ci++) overlap_(t1, t2, child(ind, ci));
            wr_raw(t1, ind, rdi(t1, child(ind, 0)) | rdi(t1, child(ind, 1)) | rdi(t1, child(ind, 2)) | rdi(t1, child(ind, 3)));
        }
    }
}
void orc_overlap(OrcTree* ground, const OrcTree* t) { Idx r = { ground->L, 1, 1 }; overlap_(ground, t, r); }
static void q_shuffle(IdxQ* q, Rng* rng) {
    int64_t n = q->n - q->head; Idx* v = q->v + q->head;
    for (int64_t i = n - 1; i > 0; i--) { int64_t j = (int64_t)(rng_u64(rng) % (uint64_t)(i + 1)); Idx t = v[i]; v[i] = v[j]; v[j] = t; }
}
static int findroom_uniform(const OrcTree* ground, IdxQ* q, Rng* rng, Idx* out) { /* qtree_functions.jl:372-402 */
    if (q_empty(q)) { Idx r = { ground->L, 1, 1 }; q_push(q, r); }
    while (!q_empty(q)) {
        Idx i = q_pop(q);
        if (i.l == 1) { if (rdi(ground, i) == ORC_EMPTY) { *out = i; return 1; } }
        else {
            int has = 0; Idx ans = i;
            const int8_t* perm = PERM4[rng_below(rng, 24)];
            for (int k = 0; k < 4; k++) {
                Idx ci = child(i, perm[k]); uint8_t g = rdi(ground, ci);
                if (g == ORC_EMPTY) { if (rng_f64(rng) < 0.7) { ans = ci; has = 1; } q_push(q, ci); }
                else if (g == ORC_MIX) q_push(q, ci);
            }
            if (!q_empty(q) && q->v[q->head].l != i.l) q_shuffle(q, rng);
            if (has) { *out = ans; return 1; }
        }
    }
    return 0;
}
/* shufflesort!(q, ground, p)  qtree_functions.jl:403-409: shuffle, then a STABLE sort by the elliptic p-norm of the cell's
 * offset from the level's centre (Julia's default sort! is stable).  Integer p: repeated multiplication (exact for p = 2,
 * the default); other p: pow(). */
static double pnorm_term(double x, double p) {
    if (p == 1.0) return x;
    if (p == 2.0) return x * x;
    if (p == 3.0) return x * x * x;
    if (p == 4.0) { double y = x * x; return y * y; }
    return pow(x, p);
}
typedef struct { double key; Idx i; } KeyedIdx;
static void keyed_merge_sort(KeyedIdx* v, KeyedIdx* tmp, int64_t n) {
    if (n < 2) return;
    int64_t h = n / 2;
    keyed_merge_sort(v, tmp, h); keyed_merge_sort(v + h, tmp, n - h);
    int64_t a = 0, b = h, o = 0;
    while (a < h && b < n) tmp[o++] = (v[b].key < v[a].key) ? v[b++] : v[a++]; /* ties keep the left element first */
    while (a < h) tmp[o++] = v[a++];
    while (b < n) tmp[o++] = v[b++];
    memcpy(v, tmp, sizeof(KeyedIdx) * (size_t)n);
}
static void q_shufflesort(IdxQ* q, const OrcTree* ground, double p, Rng* rng) {
    int64_t n = q->n - q->head; Idx* v = q->v + q->head;
    double ce = (1 + (double)LV(ground, v[0].l).sz) / 2; /* size(ground[l], 1): the padded (virtual) size of the level */
    double h = LV(ground, 1).kh, w = LV(ground, 1).kw;
    q_shuffle(q, rng);
    KeyedIdx* kv = (KeyedIdx*)malloc(sizeof(KeyedIdx) * (size_t)(2 * n));
    for (int64_t k = 0; k < n; k++) {
        kv[k].i = v[k];
        kv[k].key = pnorm_term(fabs((v[k].a - ce) / h), p) + pnorm_term(fabs(v[k].b - ce) / w, p);
    }
    keyed_merge_sort(kv, kv + n, n);
    for (int64_t k = 0; k < n; k++) v[k] = kv[k].i;
    free(kv);
}
static int findroom_gathering(const OrcTree* ground, IdxQ* q, Rng* rng, int level, double p, Idx* out) { /* qtree_functions.jl:410-441 */
    if (q_empty(q)) {
        int l = ground->L - level; if (l < 1) l = 1;
        int s = LV(ground, l).sz;
        for (int i = 1; i <= s; i++) for (int j = 1; j <= s; j++) { Idx c = { l, i, j }; if (rdi(ground, c) != ORC_FULL) q_push(q, c); }
        if (!q_empty(q)) q_shufflesort(q, ground, p, rng);
    }
    while (!q_empty(q)) {
        Idx i = q_pop(q);
        if (i.l == 1) { if (rdi(ground, i) == ORC_EMPTY) { *out = i; return 1; } }
        else {
            int has = 0; Idx ans = i;
            const int8_t* perm = PERM4[rng_below(rng, 24)];
            for (int k = 0; k < 4; k++) {
                Idx ci = child(i, perm[k]); uint8_t g = rdi(ground, ci);
                if (g == ORC_EMPTY) { if (rng_f64(rng) < 0.7) { ans = ci; has = 1; } q_push(q, ci); }
                else if (g == ORC_MIX) q_push(q, ci);
            }
            if (!q_empty(q) && q->v[q->head].l != i.l) q_shufflesort(q, ground, p, rng);
            if (has) { *out = ans; return 1; }
        }
    }
    return 0;
}
int orc_place_gathering(OrcTree* const* trees, int n, uint64_t seed, int level, double p) { /* place!(...; roomfinder=findroom_gathering) */
    OrcTree* ground = orc_tree_copy(trees[0]);
    IdxQ q; q_init(&q); Rng rng = { seed }; int ok = 0;
    for (int i = 1; i < n; i++) {
        Idx ind;
        if (!findroom_gathering(ground, &q, &rng, level, p, &ind)) { ok = -1; break; }
        int64_t ca, cb; indexcenter(ind.l, ind.a, ind.b, &ca, &cb);
        orc_setshift(trees[i], 1, (int)ca - LV(trees[i], 1).kh / 2, (int)cb - LV(trees[i], 1).kw / 2);
        orc_overlap(ground, trees[i]);
    }
    q_free(&q); orc_tree_free(ground);
    return ok;
}
int orc_place(OrcTree* const* trees, int n, uint64_t seed) { /* Stuffing.jl:64-68 + qtree_functions.jl:503-543 */
    OrcTree* ground = orc_tree_copy(trees[0]);
    IdxQ q; q_init(&q); Rng rng = { seed }; int ok = 0;
    for (int i = 1; i < n; i++) {
        Idx ind;
        if (!findroom_uniform(ground, &q, &rng, &ind)) { ok = -1; break; }
        int64_t ca, cb; indexcenter(ind.l, ind.a, ind.b, &ca, &cb);               /* setcenter!(qtree, getcenter(ind)) */
        orc_setshift(trees[i], 1, (int)ca - LV(trees[i], 1).kh / 2, (int)cb - LV(trees[i], 1).kw / 2);
        orc_overlap(ground, trees[i]);
    }
    q_free(&q); orc_tree_free(ground);
    return ok;
}
