"""Full-size GPU tests of the other BASELINE.json configurations (C4: 100k masks on a 16384^2 canvas,
C5: 4096 independent 512^2 scenes in one batch): bit-exact against the oracle where the oracle finishes in
seconds, size-independent properties at the full size."""
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
    n_scenes, per = 4096, 128
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
    sample = [0, 1, 777, 2048, 4095]
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
    s = 777
    qo = oracle_scene(s)
    lo = s * stride
    for k in range(stride):
        for l in (2, 3, 5, 10):
            assert np.array_equal(d[lo + k].level(l), qo[k].level(l)) and d[lo + k].info(l) == qo[k].info(l)
