/* stuffing_oracle.h — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the collision-and-fit hot path of guo-yong-zhi/Stuffing.jl
 * v0.10.4 (pure Julia; citations are to /root/reference/src/...).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; nothing under stuffing.jl_b200/ links or calls it.
 *
 * PARITY PINNING.  Julia is not installed in the build image and the reference has no
 * C sources, so the reference itself cannot be executed here (oracle/_ref is
 * unbuildable).  The oracle is pinned against every known-answer test the
 * reference's own test-suite holds for this path (tests/test_oracle_golden.py lists
 * them with file:line).  Everything that depends on Julia's RNG stream (child order in
 * _collision_randbfs, rand() in step!/move!, shuffle!) is "parity unpinned" against a
 * real Julia run and pinned only in CANONICAL MODE: identity child permutation,
 * list-order batchsteps, counter-based draws (include/stuffing_b200_rng.h).
 *
 * SAFE MODE.  Every pyramid read is PaddedMat's bounds-checked getindex
 * (qtrees.jl:156-162): `default` outside the kernel.  That is what the reference's CI
 * (`--check-bounds=yes`) executes through `@boundscheck return t[l][r, c]`
 * (qtrees.jl:211-215), and it equals the raw read whenever the raw read stays inside
 * the (kh+3)x(kw+3) kernel array.
 *
 * Conventions: labels and (l,a,b) are 1-based exactly as the reference prints them;
 * matrices are column-major (row index fastest).
 */
#ifndef STUFFING_ORACLE_H
#define STUFFING_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_EMPTY 1
#define ORC_FULL 2
#define ORC_MIX 3

typedef struct OrcTree OrcTree;     /* ShiftedQTree / U8SQTree   qtrees.jl:170-209 */
typedef struct OrcLinked OrcLinked; /* LinkedSpacialQTree        qtrees.jl:424-552 */
typedef struct OrcOpt OrcOpt;       /* Momentum / SGD / default  fit_helper.jl:1-38 */

/* CoItem = (i1,i2) => (l,a,b)   qtree_functions.jl:52 */
typedef struct { int32_t i1, i2, l, a, b; } OrcItem;
/* one spatial-tree record: label sits in cell (l,a,b) */
typedef struct { int32_t l, a, b, label; } OrcCellRec;

/* ---- trees ---- */
OrcTree* orc_tree_new(const uint8_t* pic, int kh, int kw, int pitch, int canvas, int dflt);
OrcTree* orc_tree_copy(const OrcTree* t);
void orc_tree_free(OrcTree* t);
int orc_tree_length(const OrcTree* t);
int orc_tree_sizelevel(const OrcTree* t);
int orc_tree_default(const OrcTree* t);
void orc_tree_level_info(const OrcTree* t, int l, int32_t out[5]); /* kh,kw,rs,cs,canvas */
void orc_tree_level_data(const OrcTree* t, int l, uint8_t* out);   /* kh*kw col-major payload */
int orc_tree_get(const OrcTree* t, int l, int a, int b);           /* safe read */
int orc_tree_set(OrcTree* t, int l, int a, int b, int v);          /* -1 = BoundsError */
void orc_level_setshift(OrcTree* t, int l, int rs, int cs);        /* PaddedMat setrshift!/setcshift! */
void orc_build(OrcTree* t, int layer);                             /* buildqtree! */
void orc_shift(OrcTree* t, int l, int s1, int s2);                 /* shift! */
void orc_setshift(OrcTree* t, int l, int s1, int s2);              /* setshift! */
void orc_shift_axis(OrcTree* t, int l, int st, int axis, int set); /* rshift!/cshift!/setrshift!/setcshift! */
/* setshift! / getshift over a label list in one call (labels == NULL: trees 1..nl); skip_same leaves a tree that already
 * sits at the requested place alone (exact); returns the number of trees rebuilt */
int64_t orc_setshifts(OrcTree* const* trees, const int32_t* labels, int nl, int level, const int32_t* rs, const int32_t* cs, int skip_same, int nthreads);
void orc_getshifts(OrcTree* const* trees, const int32_t* labels, int nl, int level, int32_t* rs, int32_t* cs);
int orc_testqtree(const OrcTree* t);                               /* utils.jl:7-24; 0 ok */
int orc_inkernelbounds(const OrcTree* t, int l, int a, int b);

/* ---- pair collision ---- */
void orc_collision_dfs(const OrcTree* q1, const OrcTree* q2, int l, int a, int b, int32_t out[3]);
/* perm_seed==0: canonical identity child order; else a seeded random order per popped node.
 * visited (may be NULL) receives the number of child nodes tested. */
void orc_collision_bfs(const OrcTree* q1, const OrcTree* q2, int l, int a, int b,
                       uint64_t perm_seed, int32_t out[3], int64_t* visited);
void orc_collision(const OrcTree* q1, const OrcTree* q2, uint64_t perm_seed, int32_t out[3]);

/* ---- many-body ---- */
/* all return the number of results, or -(needed) when cap is too small */
int64_t orc_collide_items(OrcTree* const* trees, const OrcItem* items, int64_t n, uint64_t perm_seed,
                          int nthreads, OrcItem* out, int64_t cap, int64_t* visited);
int64_t orc_totalcollisions_native(OrcTree* const* trees, int n, const int32_t* labels, int nl,
                                   uint64_t perm_seed, OrcItem* out, int64_t cap);
int64_t orc_locate_hash(OrcTree* const* trees, int n, const int32_t* labels, int nl,
                        OrcCellRec* out, int64_t cap);
/* labels==NULL: all trees.  items_out/direct_out (may be NULL) receive the narrow-phase
 * item list and the default-FULL direct hits (qtree_functions.jl:203-223). */
int64_t orc_totalcollisions_spacial(OrcTree* const* trees, int n, const int32_t* labels, int nl,
                                    int unique, uint64_t perm_seed, int nthreads,
                                    OrcItem* out, int64_t cap,
                                    OrcItem* items_out, int64_t items_cap, int64_t* n_items,
                                    int64_t* n_direct, int64_t* visited);
void orc_sort_unique(OrcItem* r, int64_t* n, int unique);

OrcLinked* orc_linked_new(int L); /* root (L,1,1); L<=0: tree without root index */
void orc_linked_free(OrcLinked* t);
void orc_linked_locate(OrcLinked* t, OrcTree* const* trees, const int32_t* labels, int nl);
void orc_linked_clear(OrcLinked* t, int label);
int64_t orc_linked_collect(const OrcLinked* t, OrcCellRec* out, int64_t cap);
int64_t orc_partialcollisions(OrcTree* const* trees, int n, OrcLinked* t, const int32_t* labels, int nl,
                              int unique, uint64_t perm_seed, int nthreads, OrcItem* out, int64_t cap,
                              OrcItem* items_out, int64_t items_cap, int64_t* n_items, int64_t* n_direct);

/* ---- fit step ---- */
void orc_grad2d(const OrcTree* t, int l, int a, int b, int32_t out[2]);
int orc_intlog2(double x);
OrcOpt* orc_opt_new(int kind /*0 default Δ/6, 1 SGD, 2 Momentum*/, double eta, double rho, int n);
void orc_opt_free(OrcOpt* o);
void orc_opt_set_eta(OrcOpt* o, double eta);
void orc_opt_reset(OrcOpt* o, int label);
void orc_opt_velocity(const OrcOpt* o, int label, double out[2], int32_t* has);
/* batchsteps! in list order (canonical mode), draws keyed (seed, iter, k, slot) */
void orc_batchsteps(OrcTree* const* trees, int n, const OrcItem* colist, int64_t nh, OrcOpt* opt,
                    uint64_t seed, uint64_t iter);
/* same with an explicit mask flag per label (batches of scenes); is_mask == NULL: label 1 */
void orc_batchsteps_masks(OrcTree* const* trees, int n, const OrcItem* colist, int64_t nh, OrcOpt* opt,
                          uint64_t seed, uint64_t iter, const uint8_t* is_mask);

/* ---- input generation (not on the hot path) ---- */
/* place!(qtrees) of Stuffing.jl:64-68: trees[0] is the mask; returns 0, or -1 "no room" */
int orc_place(OrcTree* const* trees, int n, uint64_t seed);
/* place!(...; roomfinder=findroom_gathering, level, p)  qtree_functions.jl:403-441 */
int orc_place_gathering(OrcTree* const* trees, int n, uint64_t seed, int level, double p);
void orc_overlap(OrcTree* ground, const OrcTree* t); /* overlap!(tree1, tree2) */

#ifdef __cplusplus
}
#endif
#endif
