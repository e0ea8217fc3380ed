"""Seeded synthetic inputs shared by tests/ and bench.py.  Pure numpy; no oracle, no GPU.

G-rect : randobject(sz) of /root/reference/src/utils.jl:26-30 — solid rectangles, sorted by area, largest first.
G-word : word-like binary blobs imitating examples/collision_benchmark.jl:32 (WordCloud.rendertext is not
         available): height round(6*Exp(1)+8) px, 1-8 glyph boxes of random 5x7 bit patterns scaled to that
         height, rotated by a uniform angle in [-90, 90] degrees, 1 px border.
"""
from __future__ import annotations

import math

import numpy as np

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
