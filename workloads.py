"""Seeded synthetic inputs shared by tests/ and bench.py.  Pure numpy; no oracle, no GPU.

G-rect : randobject(sz) of /root/reference/src/utils.jl:26-30 — solid rectangles, sorted by area, largest first.
G-word : word-like binary blobs imitating examples/collision_benchmark.jl:32 (WordCloud.rendertext is not
         available): height round(6*Exp(1)+8) px, 1-8 glyph boxes of random 5x7 bit patterns scaled to that
         height, rotated by a uniform angle in [-90, 90] degrees, 1 px border.
"""
from __future__ import annotations

import math
import os

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
# committed place! layouts + oracle counts (tools/make_bench_fixture.py)
FIXTURES = {"c1": "c1_word_n10000_seed1.npz", "c1k": "c1_word_n1000_seed1.npz", "c2": "c2_rect_n1000_seed2.npz", "c4": "c4_rect_n100000_seed4.npz"}

EMPTY, FULL, MIX = 1, 2, 3


def rect_objects(n, sz, rng):
    objs = []
    for _ in range(n):
        s1 = max(2, sz / 10) + 0.9 * rng.random() * sz
        s2 = max(2, sz / 10) + 0.9 * rng.random() * sz
        objs.append(np.ones((int(round(s1)), int(round(s2))), bool))
    objs.sort(key=lambda o: -o.size)
    return objs


def _rotate(img, deg):
    """nearest-neighbour rotation of a boolean image about its centre, output sized to fit"""
    th = math.radians(deg)
    c, s = math.cos(th), math.sin(th)
    h, w = img.shape
    H = int(math.ceil(abs(h * c) + abs(w * s))) + 2
    W = int(math.ceil(abs(w * c) + abs(h * s))) + 2
    yy, xx = np.mgrid[0:H, 0:W]
    y0, x0 = yy - (H - 1) / 2, xx - (W - 1) / 2
    ys = np.rint(c * y0 + s * x0 + (h - 1) / 2).astype(int)
    xs = np.rint(-s * y0 + c * x0 + (w - 1) / 2).astype(int)
    ok = (ys >= 0) & (ys < h) & (xs >= 0) & (xs < w)
    out = np.zeros((H, W), bool)
    out[ok] = img[ys[ok], xs[ok]]
    return out


def word_object(rng):
    fontsize = 6 * rng.exponential(1.0) + 8
    height = max(4, int(round(0.75 * fontsize)))  # glyph (cap) height of a font of that size
    nchar = int(rng.integers(1, 9))
    gw = max(3, int(round(0.6 * height)))
    cols = []
    for _ in range(nchar):
        g = rng.random((7, 5)) < 0.6
        g[rng.integers(0, 7), :] = True  # every glyph has a stroke
        ry = (np.arange(height) * 7 // height).clip(0, 6)
        rx = (np.arange(gw) * 5 // gw).clip(0, 4)
        cols.append(g[np.ix_(ry, rx)])
        cols.append(np.zeros((height, 1), bool))
    img = np.concatenate(cols[:-1], axis=1)
    img = _rotate(img, float(rng.integers(-90, 91)))
    rows, cs = np.flatnonzero(img.any(1)), np.flatnonzero(img.any(0))
    img = img[rows[0]:rows[-1] + 1, cs[0]:cs[-1] + 1]
    return np.pad(img, 1)  # border = 1


def word_objects(n, rng, pool=5000):
    """n word objects drawn like collision_benchmark.jl:26: a pool of `pool` words repeated to length n"""
    base = [word_object(rng) for _ in range(min(n, pool))]
    return [base[i % len(base)] for i in range(n)]


def canvas_for_mask(ms):
    """maskqtree's canvas size (Stuffing.jl:30-32)"""
    b = max(ms * 0.024, 20)
    return 2 ** math.ceil(math.log2(ms + b))


def uniform_shifts(objs, mask_side, canvas, rng):
    """uniformly random level-1 shifts that keep every object inside a centred mask_side^2 mask"""
    off = (canvas - mask_side) // 2
    rs = np.array([off + rng.integers(0, max(1, mask_side - o.shape[0])) for o in objs], np.int32)
    cs = np.array([off + rng.integers(0, max(1, mask_side - o.shape[1])) for o in objs], np.int32)
    return rs, cs


def codes_of(obj):
    return np.where(obj, FULL, EMPTY).astype(np.uint8)


def load_workload(which="c1"):
    """single-scene workloads of BASELINE.json: seeded objects + the committed place! layout (tools/make_bench_fixture.py).
    c1 (also the c3 scene): 10 000 word masks, 3500^2 mask, canvas 4096; c1k: 1 000 word masks, canvas 2048;
    c2: 1 000 rect masks sz 36, 1000^2 mask, canvas 1024 (+ the golden vector of its 50 fit rounds);
    c4: 100 000 rect masks sz 30, 15900^2 mask, canvas 16384."""
    fx = np.load(os.path.join(ROOT, "bench_data", FIXTURES[which]))
    n, seed, s = int(fx["n"]), int(fx["seed"]), int(fx["mask_side"])
    rng = np.random.default_rng(seed)
    objs = rect_objects(n, 30, rng) if which == "c4" else rect_objects(n, 36, rng) if which == "c2" else word_objects(n, rng)
    W = dict(n=n, seed=seed, mask_side=s, canvas=int(fx["canvas"]), objs=objs, rs=fx["rs"].astype(np.int32), cs=fx["cs"].astype(np.int32),
             n_items=int(fx["n_items"]), n_hits=int(fx["n_hits"]), visited=int(fx["visited"]))
    for k in ("rounds", "round_items", "round_hits", "round_visited", "final_rs", "final_cs"):
        if k in fx:
            W[k] = fx[k] if fx[k].ndim else int(fx[k])
    return W


def scene_batch(S, scenes, per_scene=128, seed=5, side=490, canvas=512):
    """C5: `scenes` = iterable of scene ids; every scene = a side^2 mask + per_scene word masks at uniform positions.
    S = the product package (qtree / maskqtree host constructors).  Returns the packed upload (payload, kh, kw, rs, cs,
    dflt, is_mask, scene ids 0..len-1, canvas)."""
    pool_rng = np.random.default_rng(seed)
    pool = [S.qtree(o, canvas) for o in word_objects(2000, pool_rng)]
    mask = S.maskqtree(np.ones((side, side), bool))
    mflat = np.asfortranarray(mask.codes).ravel(order="F")
    pflat = [np.asfortranarray(t.codes).ravel(order="F") for t in pool]
    ph, pw = np.array([t.codes.shape[0] for t in pool]), np.array([t.codes.shape[1] for t in pool])
    off = (canvas - side) // 2
    chunks, kh, kw, rs, cs, dflt, is_mask, scene = [], [], [], [], [], [], [], []
    for k, s in enumerate(scenes):
        rng = np.random.default_rng(50_000 + int(s))
        pick = rng.integers(0, len(pool), per_scene)
        chunks.append(mflat); kh.append(side); kw.append(side); rs.append(mask.shift1[0]); cs.append(mask.shift1[1]); dflt.append(mask.default); is_mask.append(1); scene.append(k)
        for j in pick:
            chunks.append(pflat[j]); kh.append(ph[j]); kw.append(pw[j]); dflt.append(pool[j].default); is_mask.append(0); scene.append(k)
            rs.append(off + int(rng.integers(0, max(1, side - ph[j])))); cs.append(off + int(rng.integers(0, max(1, side - pw[j]))))
    return dict(payload=np.concatenate(chunks), kh=np.array(kh, np.int32), kw=np.array(kw, np.int32), rs=np.array(rs, np.int32), cs=np.array(cs, np.int32),
                dflt=np.array(dflt, np.uint8), is_mask=np.array(is_mask, np.uint8), scene=np.array(scene, np.int32), canvas=canvas)


def device_scene(S, W, device=0):
    """the single-scene workload W on the GPU: mask (label 1) + objects at the fixture's place! layout"""
    s = W["mask_side"]
    trees = [S.maskqtree(np.ones((s, s), bool))] + [S.qtree(o, W["canvas"]) for o in W["objs"]]
    for t, r, c in zip(trees, W["rs"], W["cs"]):
        t.shift1 = (int(r), int(c))
    return S.DeviceQTrees(trees, device=device)


def oracle_scene(orc, W):
    """the same scene on the CPU oracle (test infrastructure / the reference arm of bench.py)"""
    qts = orc.TreeList(orc.qtrees(W["objs"], mask=np.ones((W["mask_side"], W["mask_side"]), bool)))
    orc.setshifts(qts, None, 1, W["rs"], W["cs"], skip_same=False)
    return qts
