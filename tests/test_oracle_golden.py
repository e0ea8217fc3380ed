"""Pins the CPU oracle against every known-answer test the reference's own suite holds for the
hot path.  File:line citations are to /root/reference/."""
import numpy as np
import pytest


def box(n, r0, r1, c0, c1):
    o = np.zeros((n, n))
    o[r0 - 1:r1, c0 - 1:c1] = 1  # Julia o[r0:r1, c0:c1] .= 1
    return o


def rand_pic(rng, h, w, vals=(0, 0, 1)):
    return rng.choice(np.array(vals), size=(h, w)).astype(np.int64)


def pairset(cl):
    return {frozenset(p) for p, _ in cl}


# ---- test/test_qtrees.jl:6-9 index algebra is exercised inside the linked tree (childnumber/child/parent)

def test_paddedmat_default_read_and_write(orc):
    # test/test_qtrees.jl:11-13
    rng = np.random.default_rng(0)
    qt = orc.qtree(rand_pic(rng, 120, 77), background=0)
    assert qt.get(1, -10, -15) == orc.EMPTY
    with pytest.raises(IndexError):
        qt.set(1, -10, -15, orc.EMPTY)


def test_pyramid_invariant_after_shifts(orc):
    # test/test_qtrees.jl:14-21 with testqtree (src/utils.jl:7-24)
    rng = np.random.default_rng(1)
    h, w = 93, 211
    qt = orc.qtree(rand_pic(rng, h, w), background=0)
    qt2 = orc.qtree(rand_pic(rng, h, w), background=0)
    assert qt.testqtree() == 0
    qt.shift(3, 2, 5)
    qt2.setshift(4, 1, 2)
    assert qt.testqtree() == 0
    assert qt2.testqtree() == 0
    orc.overlap(qt2, qt)
    assert qt2.testqtree() == 0


def test_degenerate_sizes(orc):
    # test/test_qtrees.jl:22-28
    assert len(orc.qtree(np.array([[1]]), background=0)) == 2
    assert len(orc.qtree(np.array([[1, 0]]), background=0)) == 2
    rng = np.random.default_rng(2)
    qt = orc.qtree(rand_pic(rng, 100, 300, (0, 0, 0, 1)), 512, background=0)
    assert len(qt) == 10 and qt.testqtree() == 0


def test_level_sizes_chain(orc):
    # 100→51→26→14→8→5→3→2
    qt = orc.qtree(box(100, 30, 50, 30, 50), background=0.0)
    assert [qt.kernelsize(l)[0] for l in range(1, 9)] == [100, 51, 26, 14, 8, 5, 3, 2]
    assert len(qt) == 8


def test_three_boxes_scene(orc):
    # test/runtests.jl:36-46
    qts = orc.qtrees([box(100, 30, 50, 30, 50), box(100, 40, 80, 40, 80), box(100, 79, 90, 79, 90)], background=0.0)
    assert len(orc.totalcollisions(qts)) == 2
    assert pairset(orc.totalcollisions_native(qts)) == {frozenset((1, 2)), frozenset((2, 3))}
    sp, det = orc.totalcollisions_spacial(qts, details=True)
    assert pairset(sp) == {frozenset((1, 2)), frozenset((2, 3))}
    # derived vectors of SURVEY §8c (identity child order), re-derived here with the C oracle
    assert orc.locate_hash(qts) == {(7, 1, 1): [1], (8, 1, 1): [2], (6, 3, 3): [3]}
    assert sorted(det["items"]) == [((1, 2), (7, 1, 1)), ((3, 2), (6, 3, 3))]
    assert sp == [((1, 2), (5, 4, 4)), ((3, 2), (5, 5, 5))]
    qts[0].setshift(1, -1000, -1000)
    assert orc.collision(qts[0], qts[1]) == (-8, 1, 1)


def test_collision_example_boxes(orc):
    # examples/collision.jl:2-20 (values derived)
    o1 = box(100, 35, 50, 30, 55)
    o2 = box(100, 45, 80, 50, 80)
    o3 = box(100, 60, 90, 67, 90)
    q = orc.qtrees([o1, o2, o3], background=0.0)
    assert orc.collision(q[0], q[1]) == (3, 13, 14)
    assert orc.collision(q[0], q[2]) == (-8, 1, 1)
    assert orc.collision_dfs(q[0], q[2]) == (-7, 1, 1)
    assert orc.collision(q[1], q[2]) == (5, 5, 5)


def test_edge_cases_batch(orc):
    # test/test_qtrees.jl:116-123
    assert orc.totalcollisions_spacial([]) == []
    assert orc.totalcollisions_native([]) == []
    qt1 = orc.qtree(np.array([[1]]))
    assert orc.totalcollisions_spacial([qt1]) == []
    assert orc.totalcollisions_native([qt1]) == []
    assert orc.partialcollisions([qt1]) == []


@pytest.mark.parametrize("kind", ["empty", "nopixel"])
def test_edge_cases_no_collision(orc, kind):
    # test/test_qtrees.jl:124-139
    if kind == "empty":
        mk = lambda: orc.qtree(np.zeros((0, 0), np.int64), background=0)
    else:
        mk = lambda: orc.qtree(np.array([[1]]))  # background = pic[1] → all background
    qt1, qt2 = mk(), mk()
    assert len(qt1) >= 2
    assert orc.collision_dfs(qt1, qt2)[0] < 0
    assert orc.collision(qt1, qt2)[0] < 0
    assert orc.totalcollisions_native([qt1, qt2]) == []
    assert orc.totalcollisions_spacial([qt1, qt2]) == []
    assert orc.partialcollisions([qt1, qt2]) == []


def test_single_pixel_trees(orc):
    # test/test_qtrees.jl:140-157
    qt1 = orc.qtree(np.array([[1]]), background=0)
    qt2 = orc.qtree(np.array([[1]]), background=0)
    assert len(qt1) >= 2
    assert orc.collision_dfs(qt1, qt2) == (1, 1, 1)
    assert orc.collision(qt1, qt2) == (1, 1, 1)
    for perm_seed in (0, 1, 2, 3):
        assert orc.collision(qt1, qt2, perm_seed) == (1, 1, 1)
    assert [c for _, c in orc.totalcollisions_native([qt1, qt2])] == [(1, 1, 1)]
    assert [c for _, c in orc.totalcollisions_spacial([qt1, qt2])] == [(1, 1, 1)]
    assert [c for _, c in orc.partialcollisions([qt1, qt2])] == [(1, 1, 1)]
    spt = orc.LinkedTree([qt1, qt2]).locate([qt1, qt2])
    assert [c for _, c in orc.partialcollisions([qt1, qt2], spt, [1])] == [(1, 1, 1)]
    qt1.setshift(1, 10, -10)
    assert orc.collision_dfs(qt1, qt2)[0] < 0
    assert orc.collision(qt1, qt2)[0] < 0
    assert orc.totalcollisions_native([qt1, qt2]) == []
    assert orc.totalcollisions_spacial([qt1, qt2]) == []
    assert orc.partialcollisions([qt1, qt2]) == []


def test_ternary_trees(orc):
    # test/test_qtrees.jl:158-169 (+ derived (l,a,b) of SURVEY §8c)
    qt1 = orc.qtree(np.full((5, 6), 3, np.uint8))
    qt2 = orc.qtree(np.full((5, 6), 3, np.uint8))
    qt3 = orc.qtree(np.full((5, 6), 2, np.uint8))
    qt4 = orc.qtree(np.full((5, 6), 1, np.uint8))
    assert orc.collision(qt1, qt2) == (-2, 1, 1)
    assert orc.collision(qt1, qt3) == (3, 1, 1)
    assert orc.collision(qt1, qt4) == (-4, 1, 1)
    assert orc.collision_dfs(qt1, qt2) == (-1, 1, 1)
    assert orc.collision_dfs(qt1, qt3) == (1, 5, 6)
    assert orc.collision_dfs(qt1, qt4) == (-4, 1, 1)
    assert orc.totalcollisions_spacial([qt1, qt2]) == []


def test_intlog2(orc):
    # test/test_fit.jl:2-3
    x = np.random.default_rng(3).random(1000) * 1000
    assert sum(int(np.floor(np.log2(v))) for v in x) == sum(orc.intlog2(v) for v in x)


def ellipse_mask(m=500, n=800):
    i, j = np.mgrid[1:m + 1, 1:n + 1]
    return ((i - m / 2) ** 2 / (m / 2) ** 2 + (j - n / 2) ** 2 / (n / 2) ** 2 < 1)


def test_constructor_conventions(orc):
    # test/runtests.jl:9-34
    qt = orc.qtree(np.ones((3, 4), bool), 16)
    assert qt.get(1, 1, 1) == orc.FULL
    qt = orc.qtree(np.ones((3, 4)), background=1.0)
    assert qt.get(1, 1, 1) == orc.EMPTY
    inside = ellipse_mask()
    maskqt = orc.maskqtree(np.where(inside, 1, 0), background=0)
    assert maskqt.get(1, *maskqt.getcenter()) == orc.EMPTY
    objs = [np.ones((3, 4), bool), np.ones((1, 5), bool), np.ones((9, 9), bool), np.ones((12, 2), bool)]
    qts = orc.qtrees(objs)
    assert qts[0].info(1)[4] == qts[-1].info(1)[4] >= 12
    qts = orc.qtrees(objs, mask=np.where(inside, 1, 0), maskbackground=1)
    assert qts[1].info(1)[4] >= 800
    assert qts[0].get(1, *qts[0].getcenter()) == orc.FULL


def test_mask_default_full_branch(orc):
    # test/test_fit.jl:27-51: an object outside the mask kernel collides with the default-FULL mask
    inside = ellipse_mask()
    objs = [np.ones((51, 40), bool)]
    qts = orc.qtrees(objs, mask=np.where(inside, 1, 0), maskbackground=0)
    qts[1].setshift(1, 1000, 1000)
    res, det = orc.totalcollisions_spacial(qts[0:2], details=True)
    assert len(res) > 0 and det["n_direct"] > 0


def randqtrees(orc, n, seed, sz=100, masksize=(1000, 1000)):
    # src/utils.jl:26-39
    rng = np.random.default_rng(seed)
    objs = []
    for _ in range(n):
        s1 = max(2, sz / 10) + 0.9 * rng.random() * sz
        s2 = max(2, sz / 10) + 0.9 * rng.random() * sz
        objs.append(np.ones((int(round(s1)), int(round(s2))), bool))
    objs.sort(key=lambda o: -o.size)
    qts = orc.qtrees(objs, mask=np.ones(masksize, bool))
    return orc.place(qts, seed)


def test_hash_vs_linked_tree(orc):
    # test/test_qtrees.jl:41-52
    qts = randqtrees(orc, 200, 11)
    L = len(qts[0])
    spt2 = orc.LinkedTree(qts).locate(qts)
    h = orc.locate_hash(qts)

    def inrange(k):
        l, a, b = k
        r = 2 ** (L - l)
        return 1 <= a <= r and 1 <= b <= r

    h_ = {k: set(v) for k, v in h.items() if inrange(k)}
    assert h_ == {k: set(v) for k, v in spt2.collect().items()}
    spt2.clear(1)
    assert all(1 not in v for v in spt2.collect().values())


def test_collision_variants_agree(orc):
    # test/test_qtrees.jl:55-69
    qts = randqtrees(orc, 200, 12)
    spt2 = orc.LinkedTree(qts).locate(qts)
    cls = orc.totalcollisions_spacial(qts)
    cln = orc.totalcollisions_native(qts)
    clp = orc.partialcollisions(qts, spt2, range(1, len(qts) + 1))
    assert len(cln) > 0
    assert pairset(cls) == pairset(cln) == pairset(clp)
    moved = sorted({x for p, _ in clp for x in p})
    assert pairset(orc.partialcollisions(qts, spt2, moved)) == pairset(cln)
    assert pairset(orc.partialcollisions(qts)) == pairset(cln)
    labels = [1, 3, 6, 99]
    sp = orc.LinkedTree(qts).locate(qts)
    c1 = orc.partialcollisions(qts, sp, labels)
    c2 = {frozenset(p) for p, _ in orc.totalcollisions(qts) if set(p) & set(labels)}
    assert c2 == pairset(c1)


def test_random_child_order_keeps_level(orc):
    # hit/miss and the hit level are invariant under shuffle4() (qtrees.jl:17)
    qts = randqtrees(orc, 120, 13)
    _, det = orc.totalcollisions_spacial(qts, details=True)
    for (i1, i2), at in det["items"][:400]:
        ref = orc.collision_bfs(qts[i1 - 1], qts[i2 - 1], at)
        for s in (5, 6, 7):
            r = orc.collision_bfs(qts[i1 - 1], qts[i2 - 1], at, perm_seed=s)
            assert r[0] == ref[0]


def test_dynamic_loops(orc):
    # test/test_qtrees.jl:71-113: partial/dynamic == native after batchsteps! and manual shifts
    qts = randqtrees(orc, 200, 14)
    n = len(qts)
    moved = list(range(1, n + 1))
    sp = orc.LinkedTree(qts)
    for it in range(10):
        c1 = orc.partialcollisions(qts, sp, moved)
        for p, _ in c1:
            for x in p:
                if x not in moved:
                    moved.append(x)
        c2 = orc.totalcollisions_native(qts)
        assert len(c1) == len(c2)
        assert pairset(c1) == pairset(c2)
        orc.batchsteps(qts, c1, None, seed=7, iteration=it)
        qts[2].shift(1, 1, 1)
        qts[6].shift(1, -1, -1)
        for x in (3, 7):
            if x not in moved:
                moved.append(x)
    moved = list(range(1, n + 1))
    unlocated = []
    sp = orc.LinkedTree(qts)
    for it in range(10):
        c1 = orc.dynamiccollisions(qts, sp, moved, unlocated)
        c2 = orc.totalcollisions_native(qts)
        c3 = orc.totalcollisions(qts)
        assert len(c1) in (len(c2), len(c3))
        assert pairset(c1) in (pairset(c2), pairset(c3))
        orc.batchsteps(qts, c1, None, seed=8, iteration=it)
        qts[2].shift(1, -1, -1)
        qts[6].shift(1, 1, 1)
        for x in (3, 7):
            if x not in moved:
                moved.append(x)


def test_fit_converges(orc):
    # the statistical end-to-end property of test/runtests.jl:48-68: trainer-E-style rounds reach 0 collisions
    qts = randqtrees(orc, 60, 15, sz=60, masksize=(500, 800))
    opt = orc.Opt("momentum", 0.25, 0.5, len(qts))
    nc = -1
    for it in range(400):
        c = orc.totalcollisions(qts, unique=False)
        nc = len(c)
        if nc == 0:
            break
        orc.batchsteps(qts, c, opt, seed=3, iteration=it)
    assert nc == 0
    assert all(t.testqtree() == 0 for t in qts[1:])


def test_place_roomfinders(orc):
    """place! (Stuffing.jl:64-68) with both room finders (qtree_functions.jl:372-441): every tree lands with its centre on
    a cell that was EMPTY in the ground built so far, the result depends on the seed only, and findroom_gathering packs
    towards the centre (its queue is sorted by the distance from the centre)."""
    import workloads
    rng = np.random.default_rng(31)
    objs = workloads.rect_objects(40, 40, rng)

    def run(**kw):
        qs = orc.qtrees(objs, mask=np.ones((260, 300), bool))
        ground = qs[0].copy()
        orc.place(qs, 11, **kw)
        for t in qs[1:]:
            ca, cb = t.getcenter()
            assert ground.get(1, ca, cb) == orc.EMPTY
            orc.overlap(ground, t)
        c = np.array([t.getcenter() for t in qs[1:]], float); m = np.array(qs[0].getcenter(), float)
        return [t.getshift() for t in qs], float(np.mean(np.hypot(*(c - m).T)))

    su, du = run()
    sg, dg = run(roomfinder="gathering")
    assert run() == (su, du) and run(roomfinder="gathering") == (sg, dg)
    assert su != sg and dg < du
    run(roomfinder="gathering", level=3, p=1)
    run(roomfinder="gathering", level=20, p=4)  # level beyond the pyramid: starts from level 1
