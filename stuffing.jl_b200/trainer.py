"""Host-side mirror of the reference's Trainer module over the C ABI (/root/reference/src/fit.jl,
fit_helper.jl).  grad2d, step!/move!, the optimiser update and the rebuild run on the GPU
(stf_batchsteps / stf_fit_round); the trainers keep the reference's host control flow.

Canonical mode replaces Julia's task-local RNG by the counter-based draws of
include/stuffing_b200_rng.h, keyed (seed, iteration, k); `iteration` advances once per batchsteps.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import i32, ptr
from .qtrees import DeviceQTrees, DeviceTree, _dev, items_to_array, items_to_list, totalcollisions, DynamicColliders


class Momentum:
    """Momentum(η=1/4, ρ=0.5)  fit_helper.jl:13-27; the velocity lives on the device, per object"""
    kind = 2

    def __init__(self, eta=0.25, rho=0.5):
        self.eta, self.rho = float(eta), float(rho)


class SGD:
    """SGD(η=1/4)  fit_helper.jl:29-38"""
    kind = 1

    def __init__(self, eta=0.25):
        self.eta, self.rho = float(eta), 0.0


class DefaultStep:
    """(t, Δ) -> Δ ./ 6, the default optimiser argument of step!/batchsteps!  fit.jl:25,69"""
    kind = 0
    eta, rho = float("nan"), 0.0


def eta(o):
    return o.eta


def set_eta(o, e):  # eta!(o, e)
    o.eta = float(e)


def reset(o, qts, labels):  # reset!.(optimiser, ts[repositioned])
    d = _dev(qts)
    lb = i32(list(labels))
    d._ck(d.lib.stf_reset_velocity(d.h, len(lb), ptr(lb)))


def grad2d(t: DeviceTree, l, a, b):
    """grad2d(t, l, a, b) → (gh, gw)  fit_helper.jl:51-66"""
    d = t.owner
    out = np.zeros(2, np.int32)
    d._ck(d.lib.stf_grad2d(d.h, 1, ptr(i32([t.label])), ptr(i32([l])), ptr(i32([a])), ptr(i32([b])), ptr(out)))
    return int(out[0]), int(out[1])


def intlog2(x: float) -> int:
    """intlog2(x::Float64) = exponent bits - 1023  fit_helper.jl:74-78"""
    return int((np.float64(x).view(np.int64) >> 52) - 1023)


class StepClock:
    """(seed, iteration) of the canonical-mode draws; one tick per batchsteps! call"""

    def __init__(self, seed=0):
        self.seed, self.iteration = int(seed), 0


_default_clock = StepClock(0)


def _set_opt(d: DeviceQTrees, optimiser):
    optimiser = optimiser or DefaultStep()
    d._ck(d.lib.stf_set_optimiser(d.h, optimiser.kind, optimiser.eta if optimiser.kind else 0.0, optimiser.rho))


def batchsteps(qts, colist, optimiser=None, clock: StepClock | None = None):
    """batchsteps!(qtrees, colist, optimiser)  fit.jl:69-75 (list order instead of shuffle!)"""
    d = _dev(qts)
    clock = clock or _default_clock
    arr = items_to_array(colist)
    _set_opt(d, optimiser)
    d._ck(d.lib.stf_batchsteps(d.h, len(arr), ptr(arr), clock.seed, clock.iteration))
    clock.iteration += 1


def fit_round(qts, labels=None, optimiser=None, clock: StepClock | None = None, partial=False):
    """one device-resident round: totalcollisions(qtrees[, labels]; unique=false) → batchsteps!; returns length(colist)"""
    d = _dev(qts)
    clock = clock or _default_clock
    _set_opt(d, optimiser)
    lb = None if labels is None else i32(list(labels))
    nh = C.c_int64(0)
    d._ck(d.lib.stf_fit_round(d.h, 1 if partial else 0, 0 if lb is None else len(lb), ptr(lb), clock.seed, clock.iteration, C.byref(nh)))
    clock.iteration += 1
    return nh.value


def fit_round_moved(qts, labels=None, optimiser=None, clock: StepClock | None = None, partial=False):
    """fit_round that also returns the distinct labels of its colist, ascending (stf_fit_round_moved): the next `moved` set of
    a dynamic loop without shipping the hit list; returns (length(colist), labels)"""
    d = _dev(qts)
    clock = clock or _default_clock
    _set_opt(d, optimiser)
    lb = None if labels is None else i32(list(labels))
    nh, nm = C.c_int64(0), C.c_int32(0)
    out = np.zeros(d.n, np.int32)
    d._ck(d.lib.stf_fit_round_moved(d.h, 1 if partial else 0, 0 if lb is None else len(lb), ptr(lb), clock.seed, clock.iteration, C.byref(nh), ptr(out), len(out), C.byref(nm)))
    clock.iteration += 1
    return nh.value, out[: nm.value].copy()


def fit_round_positions(qts, labels=None, optimiser=None, clock: StepClock | None = None, partial=False, level=1):
    """fit_round followed by getshift(t, level) of every tree, read back with the round's own synchronisation;
    returns (length(colist), rs, cs)"""
    d = _dev(qts)
    clock = clock or _default_clock
    _set_opt(d, optimiser)
    lb = None if labels is None else i32(list(labels))
    nh = C.c_int64(0)
    rs, cs = np.zeros(d.n, np.int32), np.zeros(d.n, np.int32)
    d._ck(d.lib.stf_fit_round_positions(d.h, 1 if partial else 0, 0 if lb is None else len(lb), ptr(lb), clock.seed, clock.iteration, C.byref(nh),
                                        level, ptr(rs), ptr(cs)))
    clock.iteration += 1
    return nh.value, rs, cs


def fit_rounds(qts, n_rounds, labels=None, optimiser=None, clock: StepClock | None = None, partial=False, positions=False, level=1):
    """`n_rounds` device-resident rounds of totalcollisions(qtrees[, labels]; unique=false) → batchsteps! with ONE host
    synchronisation (stf_fit_rounds); returns (per-round length(colist), totals dict[, rs, cs])"""
    d = _dev(qts)
    clock = clock or _default_clock
    _set_opt(d, optimiser)
    lb = None if labels is None else i32(list(labels))
    nh = np.zeros(max(1, n_rounds), np.int64)
    tot = np.zeros(8, np.int64)
    rs = np.zeros(d.n, np.int32) if positions else None
    cs = np.zeros(d.n, np.int32) if positions else None
    d._ck(d.lib.stf_fit_rounds(d.h, 1 if partial else 0, 0 if lb is None else len(lb), ptr(lb), clock.seed, clock.iteration, n_rounds, ptr(nh), ptr(tot),
                               level if positions else 0, ptr(rs), ptr(cs)))
    clock.iteration += n_rounds
    totals = dict(zip(_lib.COUNTER_FIELDS, (int(x) for x in tot)))
    return (nh[:n_rounds], totals, rs, cs) if positions else (nh[:n_rounds], totals)


def _labels_of(colist):
    seen, out = set(), []
    for r in colist:
        for x in (int(r["i1"]), int(r["i2"])):
            if x not in seen:
                seen.add(x)
                out.append(x)
    return out


def _note(last, colist):
    """the reference's trainers share one `colist` vector (resource[:colist]); train! reads the LAST list of the epoch"""
    if last is not None:
        last["colist"] = colist
    return colist


def trainepoch_E(qts, optimiser=None, clock=None, last=None, **_kw):
    """element-wise trainer  fit.jl:85-99"""
    d = _dev(qts)
    colist = _note(last, totalcollisions(d, unique=False, raw=True))
    nc = len(colist)
    if nc == 0:
        return nc
    batchsteps(d, colist, optimiser, clock)
    inds = _labels_of(colist)
    for ni in range(1, len(d) // len(inds) + 1):
        colist = _note(last, totalcollisions(d, inds, unique=False, raw=True))
        batchsteps(d, colist, optimiser, clock)
        if ni > 8 * len(colist):
            break
    return nc


def trainepoch_E_device(qts, optimiser=None, clock=None, last=None, chunk=8, **_kw):
    """trainepoch_E! (fit.jl:85-99) with its round loop on the device (stf_fit_epoch_e): round 0, the label set `inds` of its
    colist, the round limit and the stop rule ni > 8 length(colist) are all evaluated there; the host synchronises once per
    `chunk` rounds.  When `inds` is at or below the dispatcher's threshold the remaining rounds run through the host
    dispatcher (native all-pairs path), exactly as trainepoch_E does.  Same result as trainepoch_E."""
    d = _dev(qts)
    clock = clock or _default_clock
    _set_opt(d, optimiser)
    nh = np.zeros(chunk, np.int64)
    state, inds = np.zeros(4, np.int32), np.zeros(d.n, np.int32)
    first, nc = 0, 0
    while True:
        rd = C.c_int32(0)
        d._ck(d.lib.stf_fit_epoch_e(d.h, clock.seed, clock.iteration, first, chunk, ptr(nh), C.byref(rd), ptr(state), ptr(inds), len(inds)))
        if first == 0 and rd.value > 0:
            nc = int(nh[0])
        if not (first == 0 and rd.value == 1 and nc == 0):  # (trainepoch_E returns before batchsteps! when nc == 0: no tick)
            clock.iteration += rd.value
        first += rd.value
        if state[0] != 0 or rd.value == 0:
            break
    if last is not None:
        last["colist"] = None  # the colist stays on the device; train!'s policy re-queries when it needs it
    if state[0] == 3:  # the native dispatcher path takes over for a small label set
        inds_l = inds[: state[2]].tolist()
        for ni in range(int(state[1]) + 1, int(state[3]) + 1):
            colist = _note(last, totalcollisions(d, inds_l, unique=False, raw=True))
            batchsteps(d, colist, optimiser, clock)
            if ni > 8 * len(colist):
                break
    return nc


class LRU:
    """LRU{Int} over labels  fit_helper.jl:80-102: most recently pushed first"""

    def __init__(self):
        self.order = {}
        self.tick = 0

    def push(self, v):
        self.tick += 1
        self.order[v] = self.tick

    def collect(self, firstn=None):
        r = sorted(self.order, key=lambda k: -self.order[k])
        return r if firstn is None else r[:firstn]


class _DeviceOps:
    """what a trainer needs from its trees: their number, a many-body collision call, a batchsteps! call"""

    def __init__(self, qts, optimiser, clock):
        self.d, self.optimiser, self.clock = _dev(qts), optimiser, clock
        self.n = len(self.d)

    def total(self, labels=None):
        return totalcollisions(self.d, labels, unique=False, raw=True)

    def steps(self, colist):
        batchsteps(self.d, colist, self.optimiser, self.clock)


def em_epoch(ops, widen, memory: LRU, last=None):
    """the shared shape of trainepoch_EM!/EM2!/EM3!  fit.jl:110-212: nested rounds on LRU-widened label sets.  `ops` supplies
    n, total(labels) and steps(colist) (the device by default; the parity tests run the same policy over the oracle)"""
    colist = _note(last, ops.total())
    nc = len(colist)
    if nc == 0:
        return nc
    ops.steps(colist)

    def level(depth, parent_len, inds_src):
        inds = inds_src
        if depth < len(widen):
            for x in inds:  # push!.(memory, inds): the last pushed is the most recent
                memory.push(x)
            inds = memory.collect(len(inds) * widen[depth])
        for ni in range(1, 2 * parent_len // max(1, len(inds)) + 1):
            cl = _note(last, ops.total(inds))
            ops.steps(cl)
            if ni > 2 * len(cl):
                break
            if depth + 1 <= len(widen):
                sub = _labels_of(cl)
                if sub:
                    level(depth + 1, len(inds), sub)

    level(0, ops.n, _labels_of(colist))
    return nc


def _em(qts, widen, optimiser, clock, memory: LRU, last=None):
    return em_epoch(_DeviceOps(qts, optimiser, clock or _default_clock), widen, memory, last)


def trainepoch_EM(qts, optimiser=None, clock=None, memory=None, last=None, **_kw):
    return _em(qts, [2], optimiser, clock, memory if memory is not None else LRU(), last)


def trainepoch_EM2(qts, optimiser=None, clock=None, memory=None, last=None, **_kw):
    return _em(qts, [4, 2], optimiser, clock, memory if memory is not None else LRU(), last)


def trainepoch_EM3(qts, optimiser=None, clock=None, memory=None, last=None, **_kw):
    return _em(qts, [8, 4, 2], optimiser, clock, memory if memory is not None else LRU(), last)


def trainepoch_D(qts, optimiser=None, clock=None, moved: DynamicColliders | None = None, loops=10, last=None, **_kw):
    """dynamic trainer  fit.jl:224-233"""
    d = _dev(qts)
    moved = moved if moved is not None else DynamicColliders(d)
    colist = []
    for _ in range(loops):
        colist = _note(last, moved.dynamiccollisions(unique=False, raw=True))
        if len(colist) == 0:
            return 0
        batchsteps(d, colist, optimiser, clock)
    return len(colist)


# ------------------------------------------------------------------------------------------------ pairwise trainers
def filttrain(qts, inpool, outpool, nearlevel2, optimiser=None, clock=None):
    """filttrain!(qtrees, inpool, outpool, nearlevel2; optimiser)  fit.jl:235-266: every pair of `inpool` is tested from
    the root ((L,1,1), no bounds pre-check); pairs whose result level cp[1] >= nearlevel2 go to `outpool` (misses report
    -level of the deepest node still MIX in both trees: the nearness score), hits (cp[1] > 0) are stepped.  Canonical
    mode: pool order instead of shuffle!(inpool), colist in pool order.  One stf_collision_bfs call for the whole pool."""
    d = _dev(qts)
    if inpool is not ALLPAIRS and len(inpool) == 0:
        return 0
    if inpool is ALLPAIRS:  # every pair i < j, generated on the device in the reference's order (fit.jl:279)
        near = d._call_list(d.lib.stf_filter_pairs, 0, None, int(math.ceil(nearlevel2)))
    else:
        pool = np.ascontiguousarray(np.asarray(inpool, np.int32).reshape(-1, 2))
        near = d._call_list(d.lib.stf_filter_pairs, len(pool), ptr(pool), int(math.ceil(nearlevel2)))
    if outpool is not None:
        outpool.extend(zip(near["i1"].tolist(), near["i2"].tolist()))
    colist = near[near["l"] > 0]
    batchsteps(d, colist, optimiser, clock)
    return len(colist)


class _AllPairs:
    """the pool `[(i, j) for i in 1:n for j in i+1:n]` (fit.jl:279) without enumerating it on the host"""

    def __init__(self):
        self.n = 0

    def __len__(self):
        return self.n * (self.n - 1) // 2


ALLPAIRS = _AllPairs()


def _allpairs(n):
    return [(i, j) for i in range(1, n + 1) for j in range(i + 1, n + 1)]


def trainepoch_P(qts, optimiser=None, clock=None, nearlevel=None, last=None, **_kw):
    """pairwise trainer  fit.jl:268-296"""
    d = _dev(qts)
    nearlevel = min(-1, -d.nlevels / 2 if nearlevel is None else nearlevel)
    ALLPAIRS.n = len(d)
    indpairs, nearlist, colist = ALLPAIRS, [], []
    _note(last, colist)
    nc = filttrain(d, indpairs, nearlist, nearlevel, optimiser, clock)
    if nc == 0:
        return 0
    for ni in range(1, len(indpairs) // len(nearlist) + 1):
        colist.clear()
        nc1 = filttrain(d, nearlist, colist, 0, optimiser, clock)
        if ni > 8 * nc1:
            break
        for ci in range(1, len(nearlist) // max(1, len(colist)) + 1 if colist else 1):
            nc2 = filttrain(d, colist, None, 0, optimiser, clock)
            if ci > 4 * nc2:
                break
    return nc


def trainepoch_P2(qts, optimiser=None, clock=None, nearlevel1=None, nearlevel2=None, **_kw):
    """pairwise trainer (more levels)  fit.jl:298-341"""
    d = _dev(qts)
    nearlevel1 = min(-1, -d.nlevels * 0.75 if nearlevel1 is None else nearlevel1)
    nearlevel2 = min(-1, -d.nlevels * 0.5 if nearlevel2 is None else nearlevel2)
    ALLPAIRS.n = len(d)
    indpairs, nearlist1, nearlist2, colist = ALLPAIRS, [], [], []
    _note(_kw.get("last"), colist)
    nc = filttrain(d, indpairs, nearlist1, nearlevel1, optimiser, clock)
    if nc == 0:
        return 0
    for ni1 in range(1, len(indpairs) // len(nearlist1) + 1):
        nearlist2.clear()
        nc1 = filttrain(d, nearlist1, nearlist2, nearlevel2, optimiser, clock)
        if ni1 > nc1:
            break
        for ni2 in range(1, len(nearlist1) // max(1, len(nearlist2)) + 1 if nearlist2 else 1):
            colist.clear()
            nc2 = filttrain(d, nearlist2, colist, 0, optimiser, clock)
            if nc2 == 0 or ni2 > 4 + nc2:
                break
            for ci in range(1, len(nearlist2) // max(1, len(colist)) + 1 if colist else 1):
                nc3 = filttrain(d, colist, None, 0, optimiser, clock)
                if ci > 4 * nc3:
                    break
    return nc


def levelpools(qts, levels=None):
    """levelpools(qtrees, levels=[-length(qtrees[1]):4:-3..., -1])  fit.jl:343-352: the first pool holds every pair"""
    d = _dev(qts)
    levels = list(range(-d.nlevels, -2, 4)) + [-1] if levels is None else list(levels)
    pools = [[lv, []] for lv in levels]
    pools[0][1] = _allpairs(len(d))
    return pools


def trainepoch_Px(qts, optimiser=None, clock=None, levelpools_=None, **_kw):
    """pairwise trainer (general levels)  fit.jl:354-376: a cascade of pools, each fed by the pairs that came at least as
    near as the next pool's level"""
    d = _dev(qts)
    pools = levelpools(d) if levelpools_ is None else levelpools_
    _note(_kw.get("last"), pools[-1][1] if pools else [])
    nc = 0
    if not pools:
        return nc
    outpool = pools[1][1] if len(pools) >= 2 else None
    outlevel = pools[1][0] if len(pools) >= 2 else 0
    inpool = pools[0][1]
    if not inpool:
        return nc
    for niter in range(1, max(1, len(d) // len(inpool)) + 1):
        if outpool is not None:
            outpool.clear()
        nc = filttrain(d, inpool, outpool, outlevel, optimiser, clock)
        if nc == 0 or niter > 8 * nc:
            break
        if len(pools) >= 2:
            trainepoch_Px(d, optimiser=optimiser, clock=clock, levelpools_=pools[1:])
    return nc


_DEFAULTS = {trainepoch_E: (20, 2000), trainepoch_EM: (10, 1000), trainepoch_EM2: (10, 1000), trainepoch_EM3: (10, 1000), trainepoch_D: (10, 2000),
             trainepoch_P: (10, 100), trainepoch_P2: (2, 100), trainepoch_Px: (10, 1000)}


class MonotoneIndicator:
    """MonotoneIndicator{Int}  fit_helper.jl:102-118"""

    def __init__(self):
        self.reset()

    def reset(self):
        self.min, self.age = None, 0

    def update(self, v):
        self.age += 1
        if self.min is None or v < self.min:
            self.age, self.min = 0, v


def _pairs_of(colist):
    """first.(colist): a trainer's list holds CoItems (structured rows) or bare pairs"""
    if isinstance(colist, np.ndarray):
        return [(int(r["i1"]), int(r["i2"])) for r in colist]
    return [(int(p[0][0]), int(p[0][1])) if isinstance(p[0], (tuple, list)) else (int(p[0]), int(p[1])) for p in colist]


def outofkernelbounds(qts, mask_label=1):
    """outofkernelbounds(mask, qtrees[2:end]) .+ 1  qtrees.jl:298-306: labels whose CENTRE lies outside the mask kernel"""
    d = _dev(qts)
    rs, cs = d.getshifts()
    ca, cb = rs + np.asarray(d.kh1) // 2, cs + np.asarray(d.kw1) // 2
    m = mask_label - 1
    inside = (ca > rs[m]) & (ca <= rs[m] + d.kh1[m]) & (cb > cs[m]) & (cb <= cs[m] + d.kw1[m])
    return [int(i) + 1 for i in np.flatnonzero(~inside) if i != m]


def select_coinds(qts, colist, rng, from_=lambda i: True, force=False):
    """select_coinds  fit.jl:378-418: about 1/8 of the colliding pairs (those with the largest labels first) nominate
    their larger label for re-placement if a deterministic DFS still finds them colliding.  Canonical mode: randexp from
    the sequential generator of place!."""
    d = _dev(qts)
    pairs = _pairs_of(colist)
    n = len(pairs)
    if n == 0:
        return []
    import math
    keep = [(n / 8.0 * -math.log(1.0 - rng.f64())) > k for k in range(1, n + 1)]
    order = sorted(range(n), key=lambda k: -max(pairs[k]))  # sort!(colist, by=maximum, rev=true): stable
    pairs = [pairs[k] for k in order]
    L = d.nlevels
    arr = np.zeros(n, _lib.ITEM_DT)
    arr["i1"], arr["i2"], arr["l"], arr["a"], arr["b"] = [p[0] for p in pairs], [p[1] for p in pairs], L, 1, 1
    out = np.zeros(n, _lib.ITEM_DT)
    d._ck(d.lib.stf_collision_dfs(d.h, n, ptr(arr), ptr(out)))
    hit = out["l"] >= 0
    selected = []
    for k in range(n):
        if keep[k]:
            mij = max(pairs[k])
            if mij in selected or not from_(mij - 1):
                continue
            if hit[k]:
                selected.append(mij)
    if not selected:
        for k in range(n):
            if not keep[k]:
                mij = max(pairs[k])
                if not from_(mij - 1):
                    continue
                if force or hit[k]:
                    selected.append(mij)
                    break
    return selected


def reposition(qts, colist=None, rng=None, from_=lambda i: True, force=False):
    """reposition!  fit.jl:420-436: trees whose centre left the mask kernel are re-placed first; otherwise the trees
    select_coinds nominates.  place! = ground composition on the device + room search on the host (place.py)."""
    from .place import Rng, place
    d = _dev(qts)
    rng = rng if rng is not None else Rng(0)
    out = outofkernelbounds(d)
    if out:
        place(d, out, seed=rng.u64())
        return out
    if colist is not None:
        sel = select_coinds(d, colist, rng, from_=from_, force=force)
        if sel:
            place(d, sel, seed=rng.u64())
        return sel
    return out


def randommove(qts, colist, rng):
    """randommove!  fit.jl:437-448: a random diagonal shift of the larger tree of the last (≤ 3) pairs in label order"""
    d = _dev(qts)
    rows = list(colist) if not isinstance(colist, np.ndarray) else [r for r in colist]
    pairs = _pairs_of(colist)
    # sort!(colist, by=maximum, rev=true) (fit.jl:438): `maximum` of a CoItem (a Pair) is the lexicographic maximum of its
    # two tuples (i1, i2) and (l, a, b); of a bare pair it is max(i1, i2).  Stable, like Julia's default for `by` sorts.
    if isinstance(colist, np.ndarray):
        keys = [max(pairs[k], (int(rows[k]["l"]), int(rows[k]["a"]), int(rows[k]["b"]))) for k in range(len(pairs))]
    else:
        keys = [max(p) for p in pairs]
    order = sorted(range(len(pairs)), key=lambda k: keys[k], reverse=True)
    sel = order[max(0, len(order) - 3):]
    moved = []
    for k in sel:
        lb = max(pairs[k])
        lvl = 1
        if isinstance(colist, np.ndarray):
            lvl = max(1, int(rows[k]["l"]) - 1)
        s1 = 1 if rng.u64() & 1 else -1
        s2 = 1 if rng.u64() & 1 else -1
        d.shift([lb], lvl, [s1], [s2])
        moved.append(lb)
    return moved


def train(qts, epochs=-1, trainer=trainepoch_EM2, patience=None, optimiser=None, scheduler=lambda lr: lr * 0.5 ** 0.5,
          callback=lambda ep: None, callback_pre=lambda ep: None, reposition_=True, seed=0, **kargs):
    """train!/fit!  fit.jl:450-552, policy included: epochs, the repositioning of stuck trees (reposition!, forced after
    repeats, randommove!), out-of-bounds recovery, the η scheduler on stagnation and the global early break.
    `reposition_`: bool | float in [0,1] | int | collection of indices | predicate, as the reference's `reposition`."""
    from .place import Rng
    d = _dev(qts)
    optimiser = optimiser if optimiser is not None else Momentum(0.25)
    pat0, ep0 = _DEFAULTS.get(trainer, (10, 1000))
    patience = pat0 if patience is None else patience
    epochs = ep0 if epochs < 0 else epochs
    flag = True
    if callable(reposition_):
        from_ = reposition_
    elif isinstance(reposition_, bool):
        from_, flag = (lambda i: reposition_), reposition_
    elif isinstance(reposition_, float):
        assert 0 <= reposition_ <= 1
        th = (len(d) - 1) * (1 - reposition_)
        from_ = lambda i: i >= th  # noqa: E731
    elif isinstance(reposition_, int):
        from_ = lambda i: i >= reposition_  # noqa: E731
    else:
        rset = set(reposition_)
        from_ = lambda i: i in rset  # noqa: E731
    clock, rng = StepClock(seed), Rng(seed ^ 0x5EED)
    res, last = {}, {"colist": []}
    if trainer in (trainepoch_EM, trainepoch_EM2, trainepoch_EM3):
        res["memory"] = LRU()
    if trainer is trainepoch_D:
        res["moved"] = DynamicColliders(d)
    if trainer is trainepoch_Px:
        res["levelpools_"] = levelpools(d)
    moved = res.get("moved")
    indi_r, indi_g, indi_s = MonotoneIndicator(), MonotoneIndicator(), MonotoneIndicator()
    eta_list = []
    reposition_count, last_repositioned = 0, None
    ep = nc = 0
    while ep < epochs:
        callback_pre(ep)
        nc = trainer(d, optimiser=optimiser, clock=clock, last=last, **res, **kargs)
        colist = last["colist"] if last["colist"] is not None else []
        ep += 1
        indi_r.update(nc); indi_g.update(nc); indi_s.update(nc)
        if nc != 0 and flag and len(d) / 20 > len(colist) > 0 and patience > 0 and (indi_r.age >= patience or indi_r.age > len(colist)):
            force = reposition_count >= 2
            repositioned = reposition(d, colist, rng, from_=from_, force=force)
            if repositioned:
                indi_r.reset()
                reset(optimiser, d, repositioned)
                if moved is not None:
                    moved.union(repositioned)
            rset_now = set(repositioned)
            reposition_count = reposition_count + 1 if last_repositioned == rset_now else 0
            last_repositioned = rset_now
            if reposition_count >= 1:
                if reposition_count >= 10:
                    break  # the repositioning strategy failed
                rm = randommove(d, colist, rng)
                if moved is not None:
                    moved.union(rm)
        callback(ep)
        if nc == 0:
            outlabels = reposition(d, None, rng)
            if not outlabels:
                break
            nc += len(outlabels)
            if moved is not None:
                moved.union(outlabels)
        if indi_s.age > max(1, patience, epochs / 50):
            if not eta_list or indi_s.min < eta_list[-1][1] or (indi_s.min == eta_list[-1][1] and rng.f64() > 0.5):
                eta_list.append((optimiser.eta, indi_s.min))
                set_eta(optimiser, scheduler(optimiser.eta))
            else:
                last_eta, _ = eta_list.pop()
                set_eta(optimiser, last_eta)
            indi_s.reset()
        # length(ts) / indi_g.min is Inf when the best count is 0 (Julia float division): the test never fires then (fit.jl:547)
        ratio = float("inf") if indi_g.min == 0 else len(d) / indi_g.min
        if indi_g.age > max(2, 2 * patience, epochs / 50 * max(1, ratio)):
            break
    return ep, nc


fit = train
