"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs — bit-exact for node codes, shifts, cells, items, hits and (l,a,b); velocities within 1e-12
relative (they are FP64 on both sides and in practice identical).  Run with `-m gpu` on a B200."""
import numpy as np
import pytest

from helpers import assert_same_trees, device_from_oracle, oracle, pairset, product, rect_scene
import workloads  # noqa: E402  (helpers puts the repo root on sys.path)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    return product()


@pytest.fixture(scope="module")
def orc():
    return oracle()


def box(n, r0, r1, c0, c1):
    o = np.zeros((n, n))
    o[r0 - 1:r1, c0 - 1:c1] = 1
    return o


# ---------------------------------------------------------------------------------------------- (1) build / rebuild
def test_build_parity_random_masks(S, orc):
    rng = np.random.default_rng(100)
    shapes = [(0, 0), (1, 1), (1, 2), (2, 1), (3, 4), (5, 6), (15, 16), (16, 17), (17, 15), (31, 33), (32, 32), (33, 65),
              (64, 1), (1, 64), (100, 100), (127, 129), (200, 77), (255, 257), (300, 300)]
    qo = []
    for k, (h, w) in enumerate(shapes):
        vals = [(1, 2), (1, 1, 2), (1, 2, 3), (2,)][k % 4]
        pic = rng.choice(np.array(vals, np.uint8), size=(h, w)).astype(np.uint8)
        t = orc.OTree(pic, 512, orc.FULL if k % 5 == 0 else orc.EMPTY)
        t.level_setshift(1, int(rng.integers(-700, 700)), int(rng.integers(-700, 700)))
        qo.append(t.build())
    d = device_from_oracle(S, qo, is_mask=np.zeros(len(qo), np.uint8))
    assert_same_trees(d, qo)


def test_build_parity_big_object(S, orc):
    # a mask-sized object goes through the wide per-level kernels
    i, j = np.mgrid[1:1501, 1:1301]
    inside = ((i - 750) ** 2 / 750 ** 2 + (j - 650) ** 2 / 650 ** 2 < 1)
    qo = [orc.maskqtree(inside), orc.qtree(np.ones((40, 30), bool), 2048)]
    qo[1].setshift(1, 77, -13)
    d = device_from_oracle(S, qo)
    assert_same_trees(d, qo)
    # move the big object by odd and negative amounts at several levels
    for (l, s1, s2, mode) in [(1, -3, 5, "shift"), (3, 1, -1, "shift"), (1, -1001, 999, "set"), (5, 2, 3, "set")]:
        if mode == "shift":
            qo[0].shift(l, s1, s2); d[0].shift(l, s1, s2)
        else:
            qo[0].setshift(l, s1, s2); d[0].setshift(l, s1, s2)
        assert_same_trees(d, qo)


def test_shift_family_parity(S, orc):
    rng = np.random.default_rng(101)
    qo = []
    for k in range(12):
        h, w = int(rng.integers(1, 120)), int(rng.integers(1, 120))
        pic = rng.choice(np.array([1, 1, 2], np.uint8), size=(h, w)).astype(np.uint8)
        qo.append(orc.OTree(pic, 256, orc.EMPTY).build())
    d = device_from_oracle(S, qo, is_mask=np.zeros(len(qo), np.uint8))
    L = len(qo[0])
    for step in range(40):
        lb = int(rng.integers(1, len(qo) + 1))
        l = int(rng.integers(1, L + 1))
        s1, s2 = int(rng.integers(-5, 6)), int(rng.integers(-5, 6))
        if rng.random() < 0.5:
            qo[lb - 1].shift(l, s1, s2); d[lb - 1].shift(l, s1, s2)
        else:
            qo[lb - 1].setshift(l, s1 * 7, s2 * 7); d[lb - 1].setshift(l, s1 * 7, s2 * 7)
    assert_same_trees(d, qo)
    # test/test_qtrees.jl:16-19
    qo[0].shift(3, 2, 5); d[0].shift(3, 2, 5)
    qo[1].setshift(4, 1, 2); d[1].setshift(4, 1, 2)
    assert_same_trees(d, qo, [1, 2])
    # safe reads anywhere, inside and outside the kernels (test/test_qtrees.jl:11-13)
    n = 2000
    labels = rng.integers(1, len(qo) + 1, n); ls = rng.integers(1, L + 1, n)
    a = rng.integers(-40, 300, n); b = rng.integers(-40, 300, n)
    got = d.read_nodes(labels, ls, a, b)
    want = [qo[labels[i] - 1].get(int(ls[i]), int(a[i]), int(b[i])) for i in range(n)]
    assert list(got) == want
    assert d[0].get(1, -1000, -1500) == S.EMPTY
    with pytest.raises(IndexError):
        d[0].set(1, -1000, -1500, S.EMPTY)


# ---------------------------------------------------------------------------------------------- reference known answers
def test_reference_known_answers(S):
    # test/runtests.jl:36-46
    qts = S.qtrees([box(100, 30, 50, 30, 50), box(100, 40, 80, 40, 80), box(100, 79, 90, 79, 90)], background=0.0)
    assert len(S.totalcollisions(qts)) == 2
    assert S.totalcollisions_spacial(qts) == [((1, 2), (5, 4, 4)), ((3, 2), (5, 5, 5))]
    assert S.locate(qts).tree() == {(7, 1, 1): [1], (8, 1, 1): [2], (6, 3, 3): [3]}
    qts[0].setshift((-1000, -1000))
    assert S.collision(qts[0], qts[1]) == (-8, 1, 1)
    # examples/collision.jl:2-20
    q = S.qtrees([box(100, 35, 50, 30, 55), box(100, 45, 80, 50, 80), box(100, 60, 90, 67, 90)], background=0.0)
    assert S.collision(q[0], q[1]) == (3, 13, 14)
    assert S.collision(q[0], q[2]) == (-8, 1, 1)
    assert S.collision_dfs(q[0], q[2]) == (-7, 1, 1)
    assert S.collision(q[1], q[2]) == (5, 5, 5)


def test_reference_edge_cases(S):
    # test/test_qtrees.jl:116-169
    assert S.totalcollisions_spacial([]) == []
    assert S.totalcollisions_native([]) == []
    qt1 = S.qtree(np.array([[1]]))
    assert S.totalcollisions_spacial([qt1]) == []
    assert S.totalcollisions_native([qt1]) == []
    assert S.partialcollisions([qt1]) == []
    for mk in (lambda: S.qtree(np.zeros((0, 0), np.int64), background=0), lambda: S.qtree(np.array([[1]]))):
        qt1, qt2 = mk(), mk()
        assert len(qt1) >= 2
        assert S.collision_dfs(qt1, qt2)[0] < 0
        assert S.collision(qt1, qt2)[0] < 0
        assert S.totalcollisions_native([qt1, qt2]) == []
        assert S.totalcollisions_spacial([qt1, qt2]) == []
        assert S.partialcollisions([qt1, qt2]) == []
    qt1 = S.qtree(np.array([[1]]), background=0)
    qt2 = S.qtree(np.array([[1]]), background=0)
    assert S.collision_dfs(qt1, qt2) == (1, 1, 1)
    assert S.collision(qt1, qt2) == (1, 1, 1)
    assert [c for _, c in S.totalcollisions_native([qt1, qt2])] == [(1, 1, 1)]
    assert [c for _, c in S.totalcollisions_spacial([qt1, qt2])] == [(1, 1, 1)]
    assert [c for _, c in S.partialcollisions([qt1, qt2])] == [(1, 1, 1)]
    d = S.DeviceQTrees([qt1, qt2])
    spt = S.locate(d, S.linked_spacial_qtree(d))
    assert [c for _, c in S.partialcollisions(d, spt, [1])] == [(1, 1, 1)]
    qt1.setshift((10, -10))
    assert S.collision_dfs(qt1, qt2)[0] < 0
    assert S.collision(qt1, qt2)[0] < 0
    assert S.totalcollisions_native([qt1, qt2]) == []
    assert S.totalcollisions_spacial([qt1, qt2]) == []
    assert S.partialcollisions([qt1, qt2]) == []
    t = [S.qtree(np.full((5, 6), c, np.uint8)) for c in (3, 3, 2, 1)]
    assert S.collision(t[0], t[1]) == (-2, 1, 1)
    assert S.collision(t[0], t[2]) == (3, 1, 1)
    assert S.collision(t[0], t[3]) == (-4, 1, 1)
    assert S.collision_dfs(t[0], t[1]) == (-1, 1, 1)
    assert S.collision_dfs(t[0], t[2]) == (1, 5, 6)
    assert S.collision_dfs(t[0], t[3]) == (-4, 1, 1)
    assert S.totalcollisions_spacial([t[0], t[1]]) == []
    with pytest.raises(S._lib.StuffingError):
        S.DeviceQTrees([S.HostTree(np.zeros((2, 2), np.uint8), 4)])  # code 0 is not a node code


def test_mask_default_full_branch(S, orc):
    # test/test_fit.jl:27-51
    i, j = np.mgrid[1:501, 1:801]
    inside = ((i - 250) ** 2 / 250 ** 2 + (j - 400) ** 2 / 400 ** 2 < 1)
    qo = orc.qtrees([np.ones((51, 40), bool)], mask=np.where(inside, 1, 0), maskbackground=0)
    qo[1].setshift(1, 1000, 1000)
    d = device_from_oracle(S, qo)
    want = orc.totalcollisions_spacial(qo)
    got = S.totalcollisions_spacial(d)
    assert got == want and len(got) > 0
    assert d.counters()["direct"] > 0


# ---------------------------------------------------------------------------------------------- (2)+(3) collisions
@pytest.fixture(scope="module")
def scene200(S, orc):
    qo = rect_scene(orc, 200, 21)
    return qo, device_from_oracle(S, qo)


def test_locate_parity(S, orc, scene200):
    qo, d = scene200
    assert S.locate(d).tree() == orc.locate_hash(qo)
    labels = [5, 3, 9, 150, 1, 77]
    assert S.locate(d, labels).tree() == orc.locate_hash(qo, labels)
    lt_o = orc.LinkedTree(qo).locate(qo)
    lt_d = S.locate(d, S.linked_spacial_qtree(d))
    assert {k: set(v) for k, v in lt_d.collect_tree().items()} == {k: set(v) for k, v in lt_o.collect().items()}
    lt_d.clear(1); lt_o.clear(1)
    assert {k: set(v) for k, v in lt_d.collect_tree().items()} == {k: set(v) for k, v in lt_o.collect().items()}


def test_bfs_every_item(S, orc, scene200):
    qo, d = scene200
    _, det = orc.totalcollisions_spacial(qo, details=True)
    items = det["items"]
    assert len(items) > 500
    got = S.collision_bfs_items(d, items)
    for k, ((i1, i2), at) in enumerate(items):
        want = orc.collision_bfs(qo[i1 - 1], qo[i2 - 1], at)
        assert (int(got[k]["l"]), int(got[k]["a"]), int(got[k]["b"])) == want, (k, (i1, i2), at)
    # collision_dfs on the same pairs from the root
    pairs = sorted({(i1, i2) for (i1, i2), _ in items})[:300]
    for (i1, i2) in pairs:
        assert S.collision_dfs(d[i1 - 1], d[i2 - 1]) == orc.collision_dfs(qo[i1 - 1], qo[i2 - 1])
        assert S.collision(d[i1 - 1], d[i2 - 1]) == orc.collision(qo[i1 - 1], qo[i2 - 1])


def test_spacial_native_partial_parity(S, orc, scene200):
    qo, d = scene200
    for unique in (True, False):
        want, det = orc.totalcollisions_spacial(qo, unique=unique, details=True)
        for keep in (False, True):  # fused narrow phase (default) / materialised candidate list (the reference's itemlist)
            d.keep_items(keep)
            got = S.totalcollisions_spacial(d, unique=unique)
            assert got == sorted(want)
            c = d.counters()
            assert c["items"] == len(det["items"]) and c["direct"] == det["n_direct"]
            assert c["overflow"] > 0 or c["visited"] == det["visited"]  # node tests of the sequential BFS, early exits included
            if keep:
                assert sorted(S.items_to_list(d.fetch_items())) == sorted(det["items"])
        d.keep_items(False)
    labels = [1, 3, 6, 99, 100, 150, 42]
    assert S.totalcollisions_spacial(d, labels) == sorted(orc.totalcollisions_spacial(qo, labels))
    assert S.totalcollisions_native(d) == sorted(orc.totalcollisions_native(qo))
    assert S.totalcollisions_native(d, labels) == sorted(orc.totalcollisions_native(qo, labels))
    # test/test_qtrees.jl:55-69
    cls, cln = S.totalcollisions_spacial(d), S.totalcollisions_native(d)
    spt = S.locate(d, S.linked_spacial_qtree(d))
    clp = S.partialcollisions(d, spt, range(1, len(d) + 1))
    assert pairset(cls) == pairset(cln) == pairset(clp) and len(cln) > 0
    # partial against the oracle's linked tree, item by item
    lo = orc.LinkedTree(qo).locate(qo)
    for lbs in ([1, 3, 6, 99], [200, 5, 17, 18, 19, 1], list(range(1, 201))):
        want, det = orc.partialcollisions(qo, lo, lbs, details=True)
        for keep in (False, True):
            d.keep_items(keep)
            got = S.partialcollisions(d, spt, lbs)
            assert got == sorted(want)
            if keep:
                assert sorted(S.items_to_list(d.fetch_items())) == sorted(det["items"])
        d.keep_items(False)
    c2 = {frozenset(p) for p, _ in S.totalcollisions(d) if set(p) & {1, 3, 6, 99}}
    assert c2 == pairset(S.partialcollisions(d, spt, [1, 3, 6, 99]))


def test_fused_narrow_variant_parity(S, orc, monkeypatch):
    """STF_FUSED_NARROW=1 (read at stf_create): candidates are tested inside the candidate kernel and never stored — same
    hits, same counts, same node-test count as the oracle; the item list is only there after keep_items(True)"""
    monkeypatch.setenv("STF_FUSED_NARROW", "1")
    qo = rect_scene(orc, 200, 7, sz=100)
    d = device_from_oracle(S, qo)
    monkeypatch.delenv("STF_FUSED_NARROW")
    want, det = orc.totalcollisions_spacial(qo, unique=False, details=True)
    assert S.totalcollisions_spacial(d, unique=False) == sorted(want)
    c = d.counters()
    assert c["items"] == len(det["items"]) and c["direct"] == det["n_direct"] and (c["overflow"] > 0 or c["visited"] == det["visited"])
    with pytest.raises(S._lib.StuffingError):
        d.fetch_items()
    d.keep_items(True)
    assert S.totalcollisions_spacial(d, unique=False) == sorted(want)
    assert sorted(S.items_to_list(d.fetch_items())) == sorted(det["items"])
    d.keep_items(False)
    lo = orc.LinkedTree(qo).locate(qo)
    spt = S.locate(d, S.linked_spacial_qtree(d))
    for lbs in ([1, 3, 6, 99], list(range(1, 201))):
        assert S.partialcollisions(d, spt, lbs) == sorted(orc.partialcollisions(qo, lo, lbs))
    opt = orc.Opt("momentum", 0.25, 0.5, len(qo)); clock = S.trainer.StepClock(seed=9)
    for it in range(3):
        hits = sorted(orc.totalcollisions_spacial(qo, unique=False))
        assert S.trainer.fit_round(d, None, S.Momentum(0.25, 0.5), clock) == len(hits)
        orc.batchsteps(qo, hits, opt, seed=9, iteration=it)
    assert_same_trees(d, qo)


def test_large_frontier_path(S, orc):
    # dense noise: frontiers of thousands of MIX∧MIX nodes → the CTA-per-item path
    rng = np.random.default_rng(7)
    pics = [rng.choice(np.array([0, 0, 1]), size=(300, 280)), rng.choice(np.array([0, 0, 0, 1]), size=(290, 300)),
            rng.choice(np.array([0, 1]), size=(64, 64)), np.ones((300, 300), np.int64)]
    qo = orc.qtrees(pics, size=512, background=0)
    qo[1].setshift(1, 13, 7); qo[2].setshift(1, 100, 120); qo[3].setshift(1, 30, 30)
    d = device_from_oracle(S, qo, is_mask=np.zeros(4, np.uint8))
    # thin diagonal lines never produce a FULL code above level 1: the BFS must carry wide frontiers
    n = 1500  # (the frontier buffer of the warp-per-pair kernel holds 512 nodes: the diagonals must be longer than that)
    diag1 = np.zeros((n, n), np.int64); diag2 = np.zeros((n, n), np.int64)
    idx = np.arange(n)
    diag1[idx, idx] = 1
    diag2[idx, n - 1 - idx] = 1; diag2[idx, idx] = 1
    qo2 = orc.qtrees([diag1, diag2, diag1.copy()], size=2048, background=0)
    qo2[2].setshift(1, 1, 0)
    d2 = device_from_oracle(S, qo2, is_mask=np.zeros(3, np.uint8))
    for (dd, oo) in ((d, qo), (d2, qo2)):
        m = len(oo)
        items = [((i, j), (len(oo[0]), 1, 1)) for i in range(1, m + 1) for j in range(1, m + 1) if i != j]
        got = S.collision_bfs_items(dd, items)
        for k, ((i1, i2), at) in enumerate(items):
            assert (int(got[k]["l"]), int(got[k]["a"]), int(got[k]["b"])) == orc.collision_bfs(oo[i1 - 1], oo[i2 - 1], at)
    assert d2.counters()["overflow"] > 0


def test_out_of_canvas_objects(S, orc):
    # objects pushed far outside the canvas: negative cell coordinates, hash vs linked visibility
    qo = rect_scene(orc, 60, 33, sz=60, masksize=(500, 500))
    rng = np.random.default_rng(5)
    for lb in (3, 4, 5, 6, 7, 8):
        qo[lb - 1].setshift(1, int(rng.integers(-1500, 1500)), int(rng.integers(-1500, 1500)))
    qo[8].setshift(1, -900, -900); qo[9].setshift(1, -905, -895)
    d = device_from_oracle(S, qo)
    assert_same_trees(d, qo)
    assert S.locate(d).tree() == orc.locate_hash(qo)
    assert S.totalcollisions_spacial(d) == sorted(orc.totalcollisions_spacial(qo))
    assert S.totalcollisions_spacial(d, unique=False) == sorted(orc.totalcollisions_spacial(qo, unique=False))
    lo = orc.LinkedTree(qo); ld = S.linked_spacial_qtree(d)
    assert S.partialcollisions(d, ld) == sorted(orc.partialcollisions(qo, lo))


# ---------------------------------------------------------------------------------------------- (4) grad2d + step
def test_grad2d_parity(S, orc, scene200):
    qo, d = scene200
    hits = orc.totalcollisions_spacial(qo, unique=False)
    assert hits
    for (i1, i2), (l, a, b) in hits[:200]:
        for i in (i1, i2):
            assert S.grad2d(d[i - 1], l, a, b) == orc.grad2d(qo[i - 1], l, a, b)
    t = S.trainer
    x = np.random.default_rng(3).random(1000) * 1000
    assert sum(int(np.floor(np.log2(v))) for v in x) == sum(t.intlog2(v) for v in x)  # test/test_fit.jl:2-3


@pytest.mark.parametrize("kind", ["momentum", "sgd", "default"])
def test_batchsteps_parity(S, orc, kind):
    """fixed iteration count, same hit lists, same draws: shifts, every level and velocities must agree"""
    qo = rect_scene(orc, 150, 40, sz=80, masksize=(600, 600))
    d = device_from_oracle(S, qo)
    n = len(qo)
    oo = orc.Opt(kind, 0.25, 0.5, n)
    od = {"momentum": S.Momentum(0.25, 0.5), "sgd": S.SGD(0.25), "default": None}[kind]
    clock = S.trainer.StepClock(seed=9)
    for it in range(12):
        hits = orc.totalcollisions(qo, unique=False)
        hits_d = S.totalcollisions(d, unique=False)
        assert hits_d == sorted(hits)
        if not hits:
            break
        orc.batchsteps(qo, hits_d, oo, seed=9, iteration=it)
        S.batchsteps(d, hits_d, od, clock)
        rs, cs = d.getshifts()
        assert [tuple(x) for x in zip(rs, cs)] == [t.getshift() for t in qo]
    assert_same_trees(d, qo)
    if kind == "momentum":
        v = np.zeros(2 * n); has = np.zeros(n, np.uint8)
        d._ck(d.lib.stf_get_velocity(d.h, n, None, v.ctypes.data, has.ctypes.data))
        for lb in range(1, n + 1):
            (vo, ho) = oo.velocity(lb)
            assert bool(has[lb - 1]) == ho
            if ho:
                assert np.allclose(v[2 * lb - 2:2 * lb], vo, rtol=1e-12, atol=0)


def test_fit_round_matches_host_loop(S, orc):
    """stf_fit_round (device-resident collide → step) == totalcollisions + batchsteps on the oracle"""
    qo = rect_scene(orc, 120, 41, sz=70, masksize=(500, 500))
    d = device_from_oracle(S, qo)
    oo = orc.Opt("momentum", 0.25, 0.5, len(qo))
    clock = S.trainer.StepClock(seed=4)
    for it in range(10):
        hits = sorted(orc.totalcollisions_spacial(qo, unique=False))
        if it % 2:
            nh = S.fit_round(d, optimiser=S.Momentum(0.25, 0.5), clock=clock)
            rs, cs = d.getshifts()
        else:  # the positions ride on the round's own synchronisation
            nh, rs, cs = S.trainer.fit_round_positions(d, optimiser=S.Momentum(0.25, 0.5), clock=clock)
        assert nh == len(hits)
        orc.batchsteps(qo, hits, oo, seed=4, iteration=it)
        assert [tuple(x) for x in zip(rs, cs)] == [t.getshift() for t in qo]
    assert_same_trees(d, qo)


@pytest.mark.parametrize("n,lo,hi", [(106, 5000, 5632), (115, 5632, 8192), (130, 8192, 16384), (200, 16384, 10**9)])
def test_crowded_scene_across_the_capacity_steps(S, orc, n, lo, hi):
    """n rects on top of each other: every pair collides (n(n-1)/2 hits) and every tree is stepped ~n times in one round —
    the longest dependency chains the step kernels see.  The four sizes sit on either side of the capacity steps of the
    latency path: <= 5 632 hits the single-CTA finalize of small scenes, <= 8 192 the cluster finalize (the round is re-run
    there, the scene stays on the latency path), beyond that the throughput path, from 16 384 steps on in depth order"""
    qo = rect_scene(orc, n, 3, sz=60, masksize=(500, 500), place=False)
    for k, t in enumerate(qo[1:]):
        t.setshift(1, 270 + k % 3, 270 + k % 2)
    d = device_from_oracle(S, qo)
    oo = orc.Opt("momentum", 0.25, 0.5, len(qo))
    clock = S.trainer.StepClock(seed=2)
    seen = []
    for it in range(3):
        hits = sorted(orc.totalcollisions_spacial(qo, unique=False))
        nh, rs, cs = S.trainer.fit_round_positions(d, optimiser=S.Momentum(0.25, 0.5), clock=clock)
        seen.append(nh)
        assert nh == len(hits), (it, nh, len(hits))
        orc.batchsteps(qo, hits, oo, seed=2, iteration=it)
        assert [(int(a), int(b)) for a, b in zip(rs, cs)] == [t.getshift() for t in qo], it
    assert lo < seen[0] <= hi, seen
    assert_same_trees(d, qo)


def test_dynamic_loops(S, orc):
    # test/test_qtrees.jl:71-113 on the device, checked against the oracle step by step
    qo = rect_scene(orc, 200, 14)
    d = device_from_oracle(S, qo)
    n = len(qo)
    moved = list(range(1, n + 1))
    sp = S.linked_spacial_qtree(d)
    lo = orc.LinkedTree(qo)
    clock = S.trainer.StepClock(seed=7)
    for it in range(10):
        c1 = S.partialcollisions(d, sp, moved)
        assert c1 == sorted(orc.partialcollisions(qo, lo, moved))
        for p, _ in c1:
            for x in p:
                if x not in moved:
                    moved.append(x)
        c2 = S.totalcollisions_native(d)
        assert len(c1) == len(c2) and pairset(c1) == pairset(c2)
        S.batchsteps(d, c1, None, clock)
        orc.batchsteps(qo, c1, None, seed=7, iteration=it)
        d[2].shift(1, 1, 1); d[6].shift(1, -1, -1)
        qo[2].shift(1, 1, 1); qo[6].shift(1, -1, -1)
        for x in (3, 7):
            if x not in moved:
                moved.append(x)
    assert_same_trees(d, qo)
    colliders = S.DynamicColliders(d)
    for it in range(10):
        c1 = colliders.dynamiccollisions()
        c2 = S.totalcollisions_native(d)
        c3 = S.totalcollisions(d)
        assert len(c1) in (len(c2), len(c3))
        assert pairset(c1) in (pairset(c2), pairset(c3))
        S.batchsteps(d, c1, None, clock)
        d[2].shift(1, -1, -1); d[6].shift(1, 1, 1)
        colliders.union([3, 7])


def test_trainers_converge(S, orc):
    # the statistical property of test/runtests.jl:48-68 and test/test_fit.jl:53-75: trainers reach 0 collisions
    t = S.trainer
    for trainer in (t.trainepoch_E, t.trainepoch_EM, t.trainepoch_EM2, t.trainepoch_EM3, t.trainepoch_D):
        qo = rect_scene(orc, 40, 15, sz=60, masksize=(500, 800))
        d = device_from_oracle(S, qo)
        ep, nc = S.train(d, 300, trainer=trainer, optimiser=S.Momentum(0.25), seed=2)
        assert nc == 0, (trainer.__name__, ep, nc)
        assert S.totalcollisions(d) == []


# ---------------------------------------------------------------------------------------------- multi-scene batches
def test_scene_batch_equals_single_scenes(S, orc):
    scenes = [rect_scene(orc, 30, 50 + s, sz=50, masksize=(230, 230)) for s in range(5)]
    flat = [t for sc in scenes for t in sc]
    scene_id = [s for s, sc in enumerate(scenes) for _ in sc]
    is_mask = [int(k == 0) for sc in scenes for k in range(len(sc))]
    d = device_from_oracle(S, flat, is_mask=is_mask, scene=scene_id)
    want, off = [], 0
    for sc in scenes:
        for (i1, i2), c in orc.totalcollisions_spacial(sc, unique=False):
            want.append(((i1 + off, i2 + off), c))
        off += len(sc)
    assert S.totalcollisions_spacial(d, unique=False) == sorted(want)
    clock = S.trainer.StepClock(seed=11)
    opt = orc.Opt("momentum", 0.25, 0.5, len(flat))
    for it in range(6):
        hits = S.totalcollisions_spacial(d, unique=False)
        want, off = [], 0
        for sc in scenes:
            want += [((i1 + off, i2 + off), c) for (i1, i2), c in orc.totalcollisions_spacial(sc, unique=False)]
            off += len(sc)
        assert hits == sorted(want)
        S.batchsteps(d, hits, S.Momentum(0.25, 0.5), clock)
        orc.batchsteps(flat, hits, opt, seed=11, iteration=it, is_mask=is_mask)
    assert_same_trees(d, flat)


def test_item_parallel_shards_match_single_context(S, orc):
    """SURVEY §8e-2 on one GPU: two replicas play rank 0 and rank 1 of the item-parallel path (the all-gather is a
    device copy here; NCCL itself is covered by tools/check_shard.py under torchrun); both must stay bit-identical to
    a single context running stf_fit_round, and to the oracle."""
    import ctypes as C
    import torch
    from stuffing_b200 import parallel
    qo = rect_scene(orc, 300, 11, sz=70, masksize=(600, 600))
    ref = device_from_oracle(S, qo)
    reps = [device_from_oracle(S, qo) for _ in range(2)]
    opt = orc.Opt("momentum", 0.25, 0.5, len(qo))
    for d in [ref] + reps:
        d.lib.stf_set_optimiser(d.h, 2, 0.25, 0.5)
    for it in range(4):
        nh = C.c_int64(0)
        ref._ck(ref.lib.stf_fit_round(ref.h, 0, 0, None, 5, it, C.byref(nh)))
        nloc = []
        for r, d in enumerate(reps):
            n = C.c_int64(0)
            d._ck(d.lib.stf_shard_collide(d.h, 0, 0, None, r, 2, C.byref(n)))
            nloc.append(n.value)
        assert sum(nloc) == nh.value and min(nloc) > 0
        stride = parallel.exchange_stride(max(nloc))
        recv = torch.zeros((2 * stride, 4), dtype=torch.int32, device="cuda")
        for r, d in enumerate(reps):
            d._ck(d.lib.stf_shard_pack(d.h, recv[r * stride:].data_ptr(), stride))
        host = recv.cpu().numpy()
        assert len(parallel.unpack_segments(host, 2, stride)) == nh.value      # host twin of k_shard_unpack
        for d in reps:
            n = C.c_int64(0)
            d._ck(d.lib.stf_shard_step(d.h, recv.data_ptr(), 2, stride, 5, it, C.byref(n)))
            assert n.value == nh.value
        hits = sorted(orc.totalcollisions_spacial(qo, unique=False))
        assert len(hits) == nh.value
        orc.batchsteps(qo, hits, opt, seed=5, iteration=it)
        want = np.array([t.getshift() for t in qo])
        for d in [ref] + reps:
            rs, cs = d.getshifts()
            assert np.array_equal(rs, want[:, 0]) and np.array_equal(cs, want[:, 1])
    for d in reps:
        assert_same_trees(d, qo, labels=range(1, 60))


def test_pairwise_trainers_filttrain_parity_and_convergence(S, orc):
    """filttrain!/trainepoch_P!/P2!/Px! (fit.jl:235-376).  filttrain is checked pair by pair against the
    oracle (root BFS results incl. the miss levels used as nearness score, then batchsteps on the hits); the trainers
    must converge like the reference's (test/test_fit.jl:53-75)."""
    t = S.trainer
    qo = rect_scene(orc, 30, 23, sz=70, masksize=(400, 400))
    d = device_from_oracle(S, qo)
    L = len(qo[0])
    pairs = [(i, j) for i in range(1, len(qo) + 1) for j in range(i + 1, len(qo) + 1)]
    opt = orc.Opt("momentum", 0.25, 0.5, len(qo))
    clock = t.StepClock(seed=3)
    pool = pairs
    for rnd, nearlevel in enumerate((-L / 2, 0, 0)):
        out_dev = []
        nc = t.filttrain(d, pool, out_dev, nearlevel, S.Momentum(0.25, 0.5), clock)
        res = [orc.collision_bfs(qo[i1 - 1], qo[i2 - 1], at=(L, 1, 1)) for i1, i2 in pool]
        out_orc = [p for p, r in zip(pool, res) if r[0] >= nearlevel]
        hits = [(p, r) for p, r in zip(pool, res) if r[0] >= nearlevel and r[0] > 0]
        assert out_dev == out_orc and nc == len(hits) and (rnd > 0 or nc > 0)
        orc.batchsteps(qo, hits, opt, seed=3, iteration=rnd)
        rs, cs = d.getshifts()
        want = np.array([x.getshift() for x in qo])
        assert np.array_equal(rs, want[:, 0]) and np.array_equal(cs, want[:, 1])
        pool = out_dev
        if not pool:
            break
    # the all-pairs pool generated ON THE DEVICE (stf_filter_pairs with pairs == NULL, the reference's order of fit.jl:279)
    # gives what the explicit pool gives: same near pairs in the same order, same hits, same steps
    qo = rect_scene(orc, 30, 23, sz=70, masksize=(400, 400))
    d1, d2 = device_from_oracle(S, qo), device_from_oracle(S, qo)
    t.ALLPAIRS.n = len(qo)
    out1, out2 = [], []
    c1, c2 = t.StepClock(seed=4), t.StepClock(seed=4)
    assert t.filttrain(d1, pairs, out1, -L / 2, S.Momentum(0.25, 0.5), c1) == t.filttrain(d2, t.ALLPAIRS, out2, -L / 2, S.Momentum(0.25, 0.5), c2) > 0
    assert out1 == out2 and len(out1) < len(pairs) and d2.counters()["items"] == len(pairs)
    assert all(np.array_equal(x, y) for x, y in zip(d1.getshifts(), d2.getshifts()))
    for trainer in (t.trainepoch_P, t.trainepoch_P2, t.trainepoch_Px):
        qo = rect_scene(orc, 24, 15, sz=60, masksize=(400, 600))
        d = device_from_oracle(S, qo)
        ep, nc = S.train(d, 300, trainer=trainer, optimiser=S.Momentum(0.25), seed=2)
        assert nc == 0, (trainer.__name__, ep, nc)
        assert S.totalcollisions(d) == []


def test_place_ground_composition_and_room_search(S, orc):
    """SURVEY §8 f-1.  (a) the device ground (deepcopy(mask) + overlap! of every tree) equals the oracle's sequential
    overlap!, level by level; (b) place!(qtrees) — device ground, host room search — assigns exactly the shifts of the
    oracle's place! with the same seed; (c) place!(qtrees, inds) re-places a subset on the ground of all the others."""
    rng = np.random.default_rng(77)
    objs = workloads.rect_objects(60, 50, rng)
    qo = orc.qtrees(objs, mask=np.ones((300, 420), bool))
    d = device_from_oracle(S, qo)
    # (b) full placement
    S.place(d, seed=5)
    orc.place(qo, 5)
    rs, cs = d.getshifts()
    want = np.array([t.getshift() for t in qo])
    assert np.array_equal(rs, want[:, 0]) and np.array_equal(cs, want[:, 1])
    assert_same_trees(d, qo, labels=range(1, 12))
    # (a) ground of everything, against the oracle's overlap! chain
    g = S.overlap_ground(d)
    go = qo[0].copy()
    for t in qo[1:]:
        orc.overlap(go, t)
    for l in range(1, len(go) + 1):
        kh, kw, grs, gcs, _ = go.info(l)
        codes, hrs, hcs = g.levels[l - 1]
        assert (codes.shape[0], codes.shape[1], hrs, hcs) == (kh, kw, grs, gcs)
        assert np.array_equal(codes, go.level(l)), l
    # (c) re-place three trees on the ground of the others: they land on EMPTY ground cells
    moved = [7, 21, 40]
    g_before = S.overlap_ground(d, exclude=moved)
    S.place(d, moved, seed=9)
    for lb in moved:
        t = d[lb - 1]
        ca, cb = t.getcenter()
        assert g_before.get(1, ca, cb) == S.EMPTY


def test_train_policy_reposition(S, orc):
    """the policy layer of train! (fit.jl:378-552) over the device path: out-of-bounds recovery through
    place!, select_coinds nominations, randommove!, and convergence of a crowded scene with repositioning switched on."""
    t = S.trainer
    from stuffing_b200.place import Rng
    qo = rect_scene(orc, 50, 31, sz=60, masksize=(420, 420))
    d = device_from_oracle(S, qo)
    assert t.outofkernelbounds(d) == []
    d.setshift([5, 9], 1, [-500, 3000], [-500, 40])           # centres far outside the mask kernel
    assert t.outofkernelbounds(d) == [5, 9]
    assert t.reposition(d, None, Rng(3)) == [5, 9] and t.outofkernelbounds(d) == []
    # nominations: labels of colliding pairs only, the larger label of a pair, each at most once; `force` always names one
    d.setshift(list(range(2, 12)), 1, [200] * 10, [200] * 10)  # pile ten trees up
    colist = S.totalcollisions(d, unique=False, raw=True)
    assert len(colist) > 10
    sel = t.select_coinds(d, colist, Rng(1))
    big = {max(int(r["i1"]), int(r["i2"])) for r in colist}
    assert len(sel) == len(set(sel)) and set(sel) <= big
    assert len(t.select_coinds(d, colist[:1], Rng(2), force=True)) == 1
    before = d.getshifts()
    moved = t.randommove(d, colist, Rng(4))
    after = d.getshifts()
    assert 1 <= len(moved) <= 3 and any(before[0][m - 1] != after[0][m - 1] for m in moved)
    # a crowded scene (more area than a plain fit resolves quickly): the full policy reaches 0 collisions, everything inside
    qo = rect_scene(orc, 45, 32, sz=70, masksize=(360, 360))
    d = device_from_oracle(S, qo)
    d.setshift(list(range(2, 47)), 1, [180] * 45, [180] * 45)
    ep, nc = S.train(d, 600, trainer=t.trainepoch_EM2, optimiser=S.Momentum(0.25), patience=5, seed=6)
    assert nc == 0 and S.totalcollisions(d) == [] and t.outofkernelbounds(d) == [], (ep, nc)


@pytest.mark.gpu
def test_set_shifts_leaves_unmoved_trees_alone_and_rebuilds_moved_ones(S, orc):
    """stf_set_shifts skips a tree that already sits at the requested shift (the same rasters would be written); a tree
    that moved, or any tree after a level shift was edited by hand (a bare setshift! on one level, qtrees.jl:150-156),
    is rebuilt."""
    rng = np.random.default_rng(77)
    qo = []
    for k in range(40):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        pic = rng.choice(np.array([1, 1, 2], np.uint8), size=(h, w)).astype(np.uint8)
        t = orc.OTree(pic, 256, orc.EMPTY).build()
        t.setshift(1, int(rng.integers(0, 150)), int(rng.integers(0, 150)))
        qo.append(t)
    d = device_from_oracle(S, qo, is_mask=np.zeros(len(qo), np.uint8))
    labels = list(range(1, len(qo) + 1))
    for level in (1, 3):
        rs, cs = d.getshifts(labels, level)
        d.setshift(labels, level, rs, cs)                                   # same shift at `level` (the finer levels may snap)
        for t, r, c in zip(qo, rs, cs):
            t.setshift(level, int(r), int(c))
        d.setshift(labels, level, rs, cs)                                   # and now really nothing moves
        assert_same_trees(d, qo)
        new = [(int(r) + int(rng.integers(-9, 10)) * (k % 2), int(c) + int(rng.integers(-9, 10)) * (k % 2)) for k, (r, c) in enumerate(zip(rs, cs))]
        d.setshift(labels, level, [p[0] for p in new], [p[1] for p in new])  # every other tree may move
        for t, p in zip(qo, new):
            t.setshift(level, p[0], p[1])
        assert_same_trees(d, qo)
    # the level-3 operations left the level 2-3 rasters as they were: a level-1 setshift! to the same place is NOT a no-op
    rs, cs = d.getshifts(labels, 1)
    d.setshift(labels, 1, rs, cs)
    for t, r, c in zip(qo, rs, cs):
        t.setshift(1, int(r), int(c))
    assert_same_trees(d, qo)
    # batchsteps! move trees at the hit levels; restoring the positions afterwards must give the oracle's state
    oo, od, clock = orc.Opt("momentum", 0.25, 0.5, len(qo)), S.Momentum(0.25, 0.5), S.trainer.StepClock(seed=9)
    hits = S.totalcollisions(d, unique=False)
    assert hits == sorted(orc.totalcollisions(qo, unique=False))
    if hits:
        orc.batchsteps(qo, hits, oo, seed=9, iteration=0, is_mask=np.zeros(len(qo), np.uint8))
        S.batchsteps(d, hits, od, clock)
        assert_same_trees(d, qo)
        d.setshift(labels, 1, rs, cs)
        for t, r, c in zip(qo, rs, cs):
            t.setshift(1, int(r), int(c))
        assert_same_trees(d, qo)
    # a hand-edited level shift: the next setshift! rebuilds everything again
    rs, cs = d.getshifts(labels, 1)
    d._ck(d.lib.stf_set_level_shift(d.h, 1, 3, 5, 7))
    d.setshift(labels, 1, rs, cs)
    for t, r, c in zip(qo, rs, cs):
        t.setshift(1, int(r), int(c))
    assert_same_trees(d, qo)


@pytest.mark.gpu
@pytest.mark.parametrize("level,p", [(5, 2), (3, 1), (8, 4)])
def test_place_with_findroom_gathering(S, orc, level, p):
    """place!(...; roomfinder=findroom_gathering, level, p) (qtree_functions.jl:403-441): device ground + host search give
    the oracle's shifts for the same seed — the start cells of level max(1, L - level), the centre-first shufflesort!"""
    rng = np.random.default_rng(31)
    objs = workloads.rect_objects(40, 40, rng)
    qo = orc.qtrees(objs, mask=np.ones((260, 300), bool))
    d = device_from_oracle(S, qo)
    S.place(d, seed=11, roomfinder="gathering", level=level, p=p)
    orc.place(qo, 11, roomfinder="gathering", level=level, p=p)
    rs, cs = d.getshifts()
    want = np.array([t.getshift() for t in qo])
    assert np.array_equal(rs, want[:, 0]) and np.array_equal(cs, want[:, 1])
    assert_same_trees(d, qo, labels=range(1, 8))
    # gathering packs towards the centre: the mean distance to the canvas centre is smaller than with the uniform search
    qu = orc.qtrees(objs, mask=np.ones((260, 300), bool)); orc.place(qu, 11)
    def spread(qs):
        c = np.array([t.getcenter() for t in qs[1:]], float); m = np.array(qs[0].getcenter(), float)
        return float(np.mean(np.hypot(*(c - m).T)))
    if (level, p) == (5, 2):
        assert spread(qo) < spread(qu)


@pytest.mark.gpu
def test_set_shifts_with_resident_label_list(S, orc):
    """STF_SHIFT_DEVICE_LABELS: labels, rs and cs all live in HBM; same state as the host-list call, and a bad label in
    the resident list is dropped and reported by the next synchronising call."""
    import torch
    rng = np.random.default_rng(5)
    qo = rect_scene(orc, 60, 44, sz=60, masksize=(400, 400))
    d = device_from_oracle(S, qo)
    labels = np.array(sorted(rng.choice(np.arange(2, len(qo) + 1), 25, replace=False)), np.int32)
    rs, cs = d.getshifts(list(labels), 1)
    nrs = (rs + rng.integers(-7, 8, len(labels))).astype(np.int32); ncs = (cs + rng.integers(-7, 8, len(labels))).astype(np.int32)
    tl, tr, tc = (torch.from_numpy(x).cuda() for x in (labels, nrs, ncs))
    d._ck(d.lib.stf_set_shifts(d.h, len(labels), tl.data_ptr(), 1, tr.data_ptr(), tc.data_ptr(), 0 | 2 | 4))
    for lb, r, c in zip(labels, nrs, ncs):
        qo[lb - 1].setshift(1, int(r), int(c))
    assert_same_trees(d, qo)
    # shift! (relative) through the same path
    d._ck(d.lib.stf_set_shifts(d.h, len(labels), tl.data_ptr(), 2, tr.data_ptr(), tc.data_ptr(), 1 | 2 | 4))
    for lb, r, c in zip(labels, nrs, ncs):
        qo[lb - 1].shift(2, int(r), int(c))
    assert_same_trees(d, qo)
    # a label outside 1..n: dropped, reported at the next sync, the rest of the list is applied
    bad = labels.copy(); bad[3] = len(qo) + 7
    tb = torch.from_numpy(bad).cuda()
    d._ck(d.lib.stf_set_shifts(d.h, len(bad), tb.data_ptr(), 1, tr.data_ptr(), tc.data_ptr(), 0 | 2 | 4))
    with pytest.raises(Exception):
        d.getshifts()
        S.totalcollisions(d)
    assert d.lib.stf_set_shifts(d.h, len(labels), tl.data_ptr(), 1, tr.data_ptr(), tc.data_ptr(), 0 | 4) != 0  # needs device arrays
