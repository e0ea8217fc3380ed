"""Full-size GPU tests of every BASELINE.json configuration (C1: 10k word masks on 4096^2; C2: the 1k / 1024^2 fit loop;
C3: the dynamic loop at 10k; C4: 100k masks on a 16384^2 canvas; C5: 4096 independent 512^2 scenes in one batch):
bit-exact against the oracle where the oracle finishes in seconds, size-independent properties at the full size."""
import ctypes as C

import numpy as np
import pytest

import workloads
from helpers import assert_same_trees, device_from_oracle, oracle, product

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    return product()


@pytest.fixture(scope="module")
def orc():
    return oracle()


def _sorted_hits(orc, qo, labels=None, partial_tree=None):
    if partial_tree is not None:
        h, info = orc.partialcollisions(qo, partial_tree, labels, unique=False, raw=True)
    else:
        h, info = orc.totalcollisions_spacial(qo, labels, unique=False, raw=True)
    h = h.copy()
    orc.lib().orc_sort_unique(h.ctypes.data, C.byref(C.c_int64(len(h))), 0)
    return h, info


def test_c1_10k_words_4096_canvas(S, orc):
    """the headline scene of bench.py (BASELINE configs[0]/metric): full hit list, item multiset, fixture counts, three fit
    rounds of fitted positions, and the pyramids of moved trees — all against the oracle"""
    W = workloads.load_workload("c1")
    qo = workloads.oracle_scene(orc, W)
    d = workloads.device_scene(S, W)
    assert len(d) == 10_001 and d.nlevels == 13
    want, info = orc.totalcollisions_spacial(qo, unique=False, details=True, raw=True)
    got = S.totalcollisions_spacial(d, unique=False, raw=True)
    c = d.counters()
    assert c["items"] == info["n_items"] == W["n_items"] and c["direct"] == info["n_direct"] and len(got) == W["n_hits"]
    assert c["records"] == 10_004                                   # 10 000 objects in one cell each + the mask in four
    hs, _ = _sorted_hits(orc, qo)
    assert np.array_equal(got, hs)                                  # every hit, (i1, i2) => (l, a, b), canonical order
    assert c["overflow"] > 0 or c["visited"] == info["visited"] == W["visited"]   # node tests of the sequential BFS
    d.keep_items(True)                                              # materialise the reference's itemlist instead of testing in place
    assert np.array_equal(S.totalcollisions_spacial(d, unique=False, raw=True), hs)
    items_dev = np.sort(d.fetch_items(), order=["i1", "i2", "l", "a", "b"])
    d.keep_items(False)
    items_orc = np.sort(info["items"], order=["i1", "i2", "l", "a", "b"])
    assert np.array_equal(items_dev, items_orc)                     # the narrow-phase item multiset
    # mask excluded (examples/collision_benchmark.jl:22-37), unique
    lab = np.arange(2, len(qo) + 1, dtype=np.int32)
    assert S.items_to_list(S.totalcollisions_spacial(d, lab, raw=True)) == orc.totalcollisions_spacial(qo, lab)
    # three device-resident fit rounds: hit counts and every position after each round
    opt = orc.Opt("momentum", 0.25, 0.5, len(qo))
    clock = S.trainer.StepClock(seed=1)
    for it in range(3):
        hs, _ = _sorted_hits(orc, qo)
        nh, rs, cs = S.trainer.fit_round_positions(d, None, S.Momentum(0.25, 0.5), clock)
        orc.batchsteps(qo, hs, opt, seed=1, iteration=it)
        ro, co = orc.getshifts(qo)
        assert nh == len(hs) and np.array_equal(rs, ro) and np.array_equal(cs, co), it
    moved = sorted({int(x) for x in np.concatenate([hs["i1"], hs["i2"]]) if x != 1})[:300]
    assert_same_trees(d, qo, labels=moved + [2, 3, 5000, 10_001])


def test_c2_fit_loop_1k_rects_1024_canvas(S, orc):
    """BASELINE configs[1] (examples/packing.jl): 50 fixed rounds of totalcollisions + batchsteps on 1 000 rect masks;
    per-round hit counts and fitted positions against the committed golden vector AND a live oracle run"""
    W = workloads.load_workload("c2")
    qo = workloads.oracle_scene(orc, W)
    d = workloads.device_scene(S, W)
    assert len(d) == 1001 and d.nlevels == 11
    opt = orc.Opt("momentum", 0.25, 0.5, len(qo))
    clock = S.trainer.StepClock(seed=W["seed"])
    hits_dev = []
    for k in range(W["rounds"]):
        hs, info = _sorted_hits(orc, qo)
        nh, rs, cs = S.trainer.fit_round_positions(d, None, S.Momentum(0.25, 0.5), clock)
        c = d.counters()
        assert nh == len(hs) == W["round_hits"][k] and c["items"] + c["direct"] == info["n_items"] + info["n_direct"] == W["round_items"][k], k
        orc.batchsteps(qo, hs, opt, seed=W["seed"], iteration=k)
        ro, co = orc.getshifts(qo)
        assert np.array_equal(rs, ro) and np.array_equal(cs, co), k
        hits_dev.append(nh)
    assert np.array_equal(rs, W["final_rs"]) and np.array_equal(cs, W["final_cs"])
    assert hits_dev[0] > 100 and hits_dev[-1] == 0                   # the loop converges (test/runtests.jl:48-68)
    assert_same_trees(d, qo, labels=list(range(1, 1002, 7)))


def test_c2_fit_rounds_one_sync_equals_round_by_round(S, orc):
    """stf_fit_rounds: 50 device-resident rounds with one host synchronisation == 50 calls of stf_fit_round == the golden
    vector; a fresh context starts with buffers that are too small, so the cancel-and-continue path runs as well"""
    W = workloads.load_workload("c2")
    d = workloads.device_scene(S, W)
    clock = S.trainer.StepClock(seed=W["seed"])
    nh, tot, rs, cs = S.trainer.fit_rounds(d, W["rounds"], None, S.Momentum(0.25, 0.5), clock, positions=True)
    assert nh.tolist() == W["round_hits"].tolist() and clock.iteration == W["rounds"]
    assert np.array_equal(rs, W["final_rs"]) and np.array_equal(cs, W["final_cs"])
    assert tot["items"] == int(W["round_items"].sum()) and tot["hits"] == int(W["round_hits"].sum()) and tot["visited"] == int(W["round_visited"].sum())
    # a second pass from the restored layout, in chunks, partial mode over every movable label (same collisions: :65-69)
    lab = np.arange(2, len(d) + 1, dtype=np.int32)
    d.setshift(lab, 1, W["rs"][1:], W["cs"][1:])
    d._ck(d.lib.stf_reset_velocity(d.h, 0, None))
    S.locate(d, None, S.linked_spacial_qtree(d))
    clock = S.trainer.StepClock(seed=W["seed"])
    got = []
    for chunk in (1, 7, 30, 12):
        nh, _ = S.trainer.fit_rounds(d, chunk, lab, S.Momentum(0.25, 0.5), clock, partial=True)
        got += nh.tolist()
    rs2, cs2 = d.getshifts()
    qo = workloads.oracle_scene(orc, W)
    tree_o = orc.LinkedTree(qo); tree_o.locate(qo)
    opt = orc.Opt("momentum", 0.25, 0.5, len(qo))
    want = []
    for k in range(50):
        hs, _ = _sorted_hits(orc, qo, lab, tree_o)
        want.append(len(hs))
        orc.batchsteps(qo, hs, opt, seed=W["seed"], iteration=k)
    ro, co = orc.getshifts(qo)
    assert got == want and np.array_equal(rs2, ro) and np.array_equal(cs2, co)


def _orc_trainepoch_E(orc, qo, opt, seed, it0):
    """trainepoch_E! (fit.jl:85-99) on the oracle in canonical mode, the twin of trainer.trainepoch_E"""
    def total(labels=None):  # the dispatcher (qtree_functions.jl:270-276): native below the threshold, spacial above
        h = orc.items_array(orc.totalcollisions(qo, labels, unique=False))
        orc.lib().orc_sort_unique(h.ctypes.data, C.byref(C.c_int64(len(h))), 0)
        return h
    it = it0
    hs = total()
    nc = len(hs)
    if nc == 0:
        return nc, it
    orc.batchsteps(qo, hs, opt, seed=seed, iteration=it); it += 1
    inds = S_labels(hs)
    for ni in range(1, len(qo) // len(inds) + 1):
        hs = total(inds)
        orc.batchsteps(qo, hs, opt, seed=seed, iteration=it); it += 1
        if ni > 8 * len(hs):
            break
    return nc, it


def S_labels(colist):
    seen, out = set(), []
    for r in colist:
        for x in (int(r["i1"]), int(r["i2"])):
            if x not in seen:
                seen.add(x); out.append(x)
    return out


def test_c2_trainepoch_E_policy_parity(S, orc):
    """the reference's element-wise trainer (fit.jl:85-99: total round, then rounds over the moved labels until
    ni > 8 |colist|) at the C2 size: same epochs, same collision counts, same fitted positions as the oracle"""
    W = workloads.load_workload("c2")
    qo = workloads.oracle_scene(orc, W)
    d = workloads.device_scene(S, W)
    opt = orc.Opt("momentum", 0.25, 0.5, len(qo))
    clock = S.trainer.StepClock(seed=7)
    it = 0
    for ep in range(6):
        nc_o, it = _orc_trainepoch_E(orc, qo, opt, 7, it)
        nc_d = S.trainer.trainepoch_E(d, S.Momentum(0.25, 0.5), clock)
        assert nc_d == nc_o and clock.iteration == it, (ep, nc_d, nc_o, clock.iteration, it)
        rs, cs = d.getshifts()
        ro, co = orc.getshifts(qo)
        assert np.array_equal(rs, ro) and np.array_equal(cs, co), ep
    assert nc_o < W["round_hits"][0] // 4
    # the same epochs with the round loop ON THE DEVICE (stf_fit_epoch_e: inds, round limit and stop rule evaluated there, one
    # synchronisation per chunk of rounds; small label sets fall back to the host dispatcher): same counts, rounds, positions
    for chunk in (8, 2):
        qo2 = workloads.oracle_scene(orc, W)
        d2 = workloads.device_scene(S, W)
        opt2 = orc.Opt("momentum", 0.25, 0.5, len(qo2))
        clock2 = S.trainer.StepClock(seed=7)
        it = 0
        for ep in range(6):
            nc_o, it = _orc_trainepoch_E(orc, qo2, opt2, 7, it)
            nc_d = S.trainer.trainepoch_E_device(d2, S.Momentum(0.25, 0.5), clock2, chunk=chunk)
            assert nc_d == nc_o and clock2.iteration == it, (chunk, ep, nc_d, nc_o, clock2.iteration, it)
            rs, cs = d2.getshifts()
            ro, co = orc.getshifts(qo2)
            assert np.array_equal(rs, ro) and np.array_equal(cs, co), (chunk, ep)


class _OracleOps:
    """the trainer back end over the oracle: same dispatcher (native below the threshold), canonical list order"""

    def __init__(self, orc, qo, opt, seed):
        self.orc, self.qo, self.opt, self.seed, self.it, self.n = orc, qo, opt, seed, 0, len(qo)

    def total(self, labels=None):
        h = self.orc.items_array(self.orc.totalcollisions(self.qo, labels, unique=False))
        self.orc.lib().orc_sort_unique(h.ctypes.data, C.byref(C.c_int64(len(h))), 0)
        return h

    def steps(self, colist):
        self.orc.batchsteps(self.qo, colist, self.opt, seed=self.seed, iteration=self.it)
        self.it += 1


def test_c2_default_trainer_EM2_policy_parity(S, orc):
    """the reference's DEFAULT trainer (train!'s trainepoch_EM2!, fit.jl:137-171: LRU-widened label sets x4 / x2, three
    nested round loops) at the C2 size: the same policy code runs once over the device and once over the oracle; collision
    counts, number of rounds and every position must agree epoch by epoch (label lists in LRU order, native/spacial
    dispatch included)"""
    W = workloads.load_workload("c2")
    qo = workloads.oracle_scene(orc, W)
    d = workloads.device_scene(S, W)
    t = S.trainer
    ops_o = _OracleOps(orc, qo, orc.Opt("momentum", 0.25, 0.5, len(qo)), 11)
    clock = t.StepClock(seed=11)
    mem_d, mem_o = t.LRU(), t.LRU()
    for ep in range(4):
        nc_o = t.em_epoch(ops_o, [4, 2], mem_o)
        nc_d = t.trainepoch_EM2(d, S.Momentum(0.25, 0.5), clock, memory=mem_d)
        assert nc_d == nc_o and clock.iteration == ops_o.it, (ep, nc_d, nc_o, clock.iteration, ops_o.it)
        rs, cs = d.getshifts()
        ro, co = orc.getshifts(qo)
        assert np.array_equal(rs, ro) and np.array_equal(cs, co), ep
    assert clock.iteration > 12 and nc_o < W["round_hits"][0] // 2


def test_c3_dynamic_loop_10k(S, orc):
    """BASELINE configs[2]: 20 steps of the scripted dynamic loop of bench.py --workload c3 on the 10k scene — per step the
    partialcollisions result through the linked tree and every position after batchsteps, against the oracle"""
    import bench
    W = workloads.load_workload("c1")
    qo = workloads.oracle_scene(orc, W)
    d = workloads.device_scene(S, W)
    n_all = len(d)
    tree_d = S.linked_spacial_qtree(d); S.locate(d, None, tree_d)
    tree_o = orc.LinkedTree(qo); tree_o.locate(qo)
    opt = orc.Opt("momentum", 0.25, 0.5, n_all)
    S.trainer._set_opt(d, S.Momentum(0.25, 0.5))
    moved = list(range(2, n_all + 1))
    total_hits = 0
    for step, (jit, dr, dc) in enumerate(bench.c3_script(n_all, 20)):
        d.shift(jit, 1, dr, dc)
        for lb, a, b in zip(jit.tolist(), dr.tolist(), dc.tolist()):
            qo[lb - 1].shift(1, a, b)
        labels = list(dict.fromkeys(moved + jit.tolist()))
        got = S.partialcollisions(d, tree_d, labels, unique=False, raw=True)
        want, info = _sorted_hits(orc, qo, labels, tree_o)
        c = d.counters()
        assert np.array_equal(got, want), step
        assert c["items"] == info["n_items"] and c["direct"] == info["n_direct"], step
        d._ck(d.lib.stf_batchsteps(d.h, len(got), got.ctypes.data, 1, step))
        orc.batchsteps(qo, want, opt, seed=1, iteration=step)
        rs, cs = d.getshifts()
        ro, co = orc.getshifts(qo)
        assert np.array_equal(rs, ro) and np.array_equal(cs, co), step
        moved = bench.labels_of_result(got)
        total_hits += len(got)
    assert total_hits > 5000 and len(moved) < n_all // 2
    # the same loop as ONE call per step (stf_fit_round_moved: the hit list stays on the device, the labels of the result
    # come back): same hit counts, same moved sets, same positions as the two-call loop above
    d2 = workloads.device_scene(S, W)
    S.locate(d2, None, S.linked_spacial_qtree(d2))
    clock = S.trainer.StepClock(seed=1)
    qo2 = workloads.oracle_scene(orc, W)
    tree_o2 = orc.LinkedTree(qo2); tree_o2.locate(qo2)
    opt2 = orc.Opt("momentum", 0.25, 0.5, n_all)
    moved = list(range(2, n_all + 1))
    for step, (jit, dr, dc) in enumerate(bench.c3_script(n_all, 8)):
        d2.shift(jit, 1, dr, dc)
        for lb, a, b in zip(jit.tolist(), dr.tolist(), dc.tolist()):
            qo2[lb - 1].shift(1, a, b)
        labels = list(dict.fromkeys(moved + jit.tolist()))
        nh, mv = S.trainer.fit_round_moved(d2, labels, S.Momentum(0.25, 0.5), clock, partial=True)
        want, _ = _sorted_hits(orc, qo2, labels, tree_o2)
        orc.batchsteps(qo2, want, opt2, seed=1, iteration=step)
        assert nh == len(want) and mv.tolist() == sorted(set(want["i1"].tolist()) | set(want["i2"].tolist())), step
        rs, cs = d2.getshifts(); ro, co = orc.getshifts(qo2)
        assert np.array_equal(rs, ro) and np.array_equal(cs, co), step
        moved = [x for x in mv.tolist() if x != 1]


def test_two_contexts_in_one_process(S, orc):
    """a host that drives several contexts from one process (SURVEY §8b threading row): function attributes are set per
    context/device in stf_create; two live contexts give identical, oracle-exact results (on a multi-GPU box the second one
    sits on device 1)"""
    import torch
    ndev = torch.cuda.device_count()
    W = workloads.load_workload("c1k")
    qo = workloads.oracle_scene(orc, W)
    want = sorted(orc.totalcollisions_spacial(qo, unique=False))
    ds = [workloads.device_scene(S, W, device=(k % ndev)) for k in range(2)]
    for d in ds:
        assert S.totalcollisions_spacial(d, unique=False) == want and len(want) == W["n_hits"]
    clock = [S.trainer.StepClock(seed=3), S.trainer.StepClock(seed=3)]
    for it in range(2):
        res = [S.trainer.fit_round_positions(d, None, S.Momentum(0.25, 0.5), clock[k]) for k, d in enumerate(ds)]
        assert res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])


def test_gpu_group_item_parallel_rounds(S, orc):
    """stf_group_*: one process drives several contexts; every round is split by candidate records and the hit lists are
    exchanged by the GPUs themselves (peer stores + device-side flag barrier).  Replicas must stay bit-identical to each
    other, to a single-context stf_fit_round, and to the oracle.  On a one-GPU box the group's contexts share device 0."""
    import torch
    from stuffing_b200 import parallel
    ndev = torch.cuda.device_count()
    devices = list(range(min(ndev, 4))) if ndev > 1 else [0, 0, 0]
    W = workloads.load_workload("c1k")
    qo = workloads.oracle_scene(orc, W)
    s = W["mask_side"]
    trees = [S.maskqtree(np.ones((s, s), bool))] + [S.qtree(o, W["canvas"]) for o in W["objs"]]
    for t, r, c in zip(trees, W["rs"], W["cs"]):
        t.shift1 = (int(r), int(c))
    g = parallel.GpuGroup(trees, devices)
    single = S.DeviceQTrees(trees)
    opt = orc.Opt("momentum", 0.25, 0.5, len(qo))
    g.each(lambda d: S.trainer._set_opt(d, S.Momentum(0.25, 0.5)))
    clock = S.trainer.StepClock(seed=5)
    for it in range(5):
        hs, info = _sorted_hits(orc, qo)
        nh_g = g.fit_round(5, it)
        nh_s = S.trainer.fit_round(single, None, S.Momentum(0.25, 0.5), clock)
        assert nh_g == nh_s == len(hs), (it, nh_g, nh_s, len(hs))
        assert sum(d.counters()["items"] + d.counters()["direct"] for d in g.replicas) == info["n_items"] + info["n_direct"]  # the shares add up
        orc.batchsteps(qo, hs, opt, seed=5, iteration=it)
        ro, co = orc.getshifts(qo)
        rs, cs = single.getshifts()
        assert np.array_equal(rs, ro) and np.array_equal(cs, co)
        for d in g.replicas:
            r2, c2 = d.getshifts()
            assert np.array_equal(r2, ro) and np.array_equal(c2, co), it
        assert g.nvlink_bytes == (len(devices) - 1) * (16 * nh_g + 16 * len(devices))
    # partialcollisions mode over a moved subset
    labels = list(range(2, 400))
    for d in g.replicas + [single]:
        S.locate(d, None, S.linked_spacial_qtree(d))
    nh_g = g.fit_round(5, 9, labels=labels, partial=True)
    clock9 = S.trainer.StepClock(seed=5); clock9.iteration = 9
    nh_s = S.trainer.fit_round(single, labels, S.Momentum(0.25, 0.5), clock9, partial=True)
    assert nh_g == nh_s
    rs, cs = single.getshifts()
    for d in g.replicas:
        r2, c2 = d.getshifts()
        assert np.array_equal(r2, rs) and np.array_equal(c2, cs)
    g.close()


def test_c4_100k_objects_16384_canvas(S, orc):
    rng = np.random.default_rng(4)
    n = 100_000
    objs = workloads.rect_objects(n, 30, rng)
    qo = orc.qtrees(objs, mask=np.ones((15900, 15900), bool))
    orc.place(qo, 4)
    assert len(qo[0]) == 15
    d = device_from_oracle(S, qo)
    assert_same_trees(d, qo, labels=[1, 2, 3, 50_000, 100_001])   # the 15900^2 mask goes through the wide build path
    want, info = orc.totalcollisions_spacial(qo, unique=False, raw=True)
    got = S.totalcollisions_spacial(d, unique=False, raw=True)
    c = d.counters()
    assert c["items"] == info["n_items"] and c["direct"] == info["n_direct"]
    assert S.items_to_list(got) == sorted(orc._items_out(want, len(want))) and len(got) > 100
    assert S.items_to_list(S.totalcollisions_spacial(d, raw=True)) == orc.totalcollisions_spacial(qo)
    # three fit iterations: fitted positions must agree exactly
    opt = orc.Opt("momentum", 0.25, 0.5, len(qo))
    clock = S.trainer.StepClock(seed=44)
    for it in range(3):
        hits = S.totalcollisions_spacial(d, unique=False, raw=True)
        assert S.items_to_list(hits) == sorted(orc.totalcollisions_spacial(qo, unique=False))
        S.batchsteps(d, hits, S.Momentum(0.25, 0.5), clock)
        orc.batchsteps(qo, hits, opt, seed=44, iteration=it)
        rs, cs = d.getshifts()
        ro = np.array([t.getshift() for t in qo])
        assert np.array_equal(rs, ro[:, 0]) and np.array_equal(cs, ro[:, 1])
    moved = sorted({int(x) for h in hits for x in (h["i1"], h["i2"]) if x != 1})[:200]
    assert_same_trees(d, qo, labels=moved)


def _scene_batch(S, n_scenes, per_scene, seed):
    rng = np.random.default_rng(seed)
    pool = workloads.word_objects(2000, rng)
    side, canvas = 490, 512
    trees, scene, is_mask = [], [], []
    for s in range(n_scenes):
        trees.append(S.maskqtree(np.ones((side, side), bool)))
        scene.append(s); is_mask.append(1)
        for k in range(per_scene):
            o = pool[int(rng.integers(0, len(pool)))]
            t = S.qtree(o, canvas)
            off = (canvas - side) // 2
            t.shift1 = (off + int(rng.integers(0, max(1, side - o.shape[0]))), off + int(rng.integers(0, max(1, side - o.shape[1]))))
            trees.append(t); scene.append(s); is_mask.append(0)
    return trees, np.array(scene, np.int32), np.array(is_mask, np.uint8)


def test_c5_4096_scene_batch(S, orc):
    _check_scene_batch(S, orc, 4096, [0, 1, 777, 2048, 4095], 777)


def test_c5_mid_size_batch_400_scenes(S, orc):
    """a rank's share of a split batch: ~50 k records and ~26 k hits per step — the sizes at which the one-launch radix
    passes (decoupled look-back) and the scene-by-scene step kernel are used instead of the multi-launch versions"""
    d = _check_scene_batch(S, orc, 400, [0, 1, 123, 399], 123)
    # the multi-launch versions (three launches per radix pass, 16 level sweeps) must leave exactly the same state
    import os
    os.environ["STF_SORT_3LAUNCH"] = "1"
    try:
        d3 = _check_scene_batch(S, orc, 400, [0], 0)
    finally:
        del os.environ["STF_SORT_3LAUNCH"]
    for l in (1, 3):
        a, b = d.getshifts(level=l), d3.getshifts(level=l)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # ... and so must the steps through global tickets and flags in depth order (the path of batches whose scene ids do not
    # ascend with the labels) instead of scene by scene (k_batchsteps_scene)
    os.environ["STF_STEP_GLOBAL"] = "1"
    try:
        dg = _check_scene_batch(S, orc, 400, [0], 0)
    finally:
        del os.environ["STF_STEP_GLOBAL"]
    for l in (1, 3):
        a, b = d.getshifts(level=l), dg.getshifts(level=l)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def _check_scene_batch(S, orc, n_scenes, sample, probe):
    per = 128
    trees, scene, is_mask = _scene_batch(S, n_scenes, per, 5)
    d = S.DeviceQTrees(trees, is_mask=is_mask, scene=scene)
    assert len(d) == n_scenes * (per + 1) and d.nlevels == 10
    hits = S.totalcollisions_spacial(d, unique=False, raw=True)
    assert len(hits) > n_scenes
    # (a) run-to-run determinism, (b) no pair crosses a scene boundary
    assert np.array_equal(hits, S.totalcollisions_spacial(d, unique=False, raw=True))
    assert np.array_equal(scene[hits["i1"] - 1], scene[hits["i2"] - 1])
    # (c) every hit is a FULL code of the pair at the reported node (qtree_functions.jl:31-33)
    sub = hits[:: max(1, len(hits) // 20000)]
    c1 = d.read_nodes(sub["i1"], sub["l"], sub["a"], sub["b"]); c2 = d.read_nodes(sub["i2"], sub["l"], sub["a"], sub["b"])
    assert np.all((c1 & c2) == S.FULL)
    # (d) sampled scenes against the oracle, bit for bit, before and after two batched fit iterations
    stride = per + 1

    def oracle_scene(s):
        lo = s * stride
        qo = [orc.OTree(trees[lo + k].codes, 512, trees[lo + k].default) for k in range(stride)]
        rs, cs = d.getshifts(list(range(lo + 1, lo + stride + 1)))
        for k, t in enumerate(qo):
            t.setshift(1, int(rs[k]), int(cs[k]))
        return qo

    def device_scene_hits(h, s):
        lo = s * stride
        m = (h["i1"] > lo) & (h["i1"] <= lo + stride)
        return [((int(r["i1"]) - lo, int(r["i2"]) - lo), (int(r["l"]), int(r["a"]), int(r["b"]))) for r in h[m]]

    for s in sample:
        assert device_scene_hits(hits, s) == sorted(orc.totalcollisions_spacial(oracle_scene(s), unique=False))
    clock = S.trainer.StepClock(seed=55)
    for it in range(2):
        S.batchsteps(d, hits, S.Momentum(0.25, 0.5), clock)
        hits = S.totalcollisions_spacial(d, unique=False, raw=True)
    for s in sample:
        assert device_scene_hits(hits, s) == sorted(orc.totalcollisions_spacial(oracle_scene(s), unique=False))
    # the pyramids of a moved scene are consistent with a from-scratch build at the fitted shifts
    s = probe
    qo = oracle_scene(s)
    lo = s * stride
    for k in range(stride):
        for l in (2, 3, 5, 10):
            assert np.array_equal(d[lo + k].level(l), qo[k].level(l)) and d[lo + k].info(l) == qo[k].info(l)
    return d
