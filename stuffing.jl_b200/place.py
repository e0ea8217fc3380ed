"""place! / reposition support (/root/reference/src/qtree_functions.jl:372-543, Stuffing.jl:64-73).

The work splits by shape.  Composing the ground — deepcopy(mask) with thousands of trees overlap!-ed into it — is
data-parallel and runs on the GPU (stf_ground_compose: two commutative atomics per touched word + one pyramid build).
The room search (findroom_uniform) is a sequential BFS whose order is driven by ONE random stream (shuffle4, rand() < 0.7,
shuffle!(queue)), a few hundred node reads per tree: it stays on the host, reads the composed ground, and overlaps the few
trees it places into its host copy.  Canonical mode: the sequential generator of the oracle (splitmix-style, seeded) in
place of Julia's task-local RNG, identity otherwise."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import i32, ptr
from .qtrees import EMPTY, FULL, MIX, DeviceQTrees

_M64 = (1 << 64) - 1
_G = 0x9E3779B97F4A7C15
PERM4 = [(0, 1, 2, 3), (0, 1, 3, 2), (0, 2, 1, 3), (0, 2, 3, 1), (0, 3, 1, 2), (0, 3, 2, 1), (1, 0, 2, 3), (1, 0, 3, 2), (1, 2, 0, 3), (1, 2, 3, 0), (1, 3, 0, 2), (1, 3, 2, 0),
         (2, 0, 1, 3), (2, 0, 3, 1), (2, 1, 0, 3), (2, 1, 3, 0), (2, 3, 0, 1), (2, 3, 1, 0), (3, 0, 1, 2), (3, 0, 2, 1), (3, 1, 0, 2), (3, 1, 2, 0), (3, 2, 0, 1), (3, 2, 1, 0)]  # qtrees.jl:12-15


def _mix64(x):
    x = (x + _G) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


class Rng:
    """the sequential generator of the canonical-mode place! (shared with the oracle)"""

    def __init__(self, seed):
        self.s = int(seed) & _M64

    def u64(self):
        self.s = (self.s + _G) & _M64
        return _mix64(self.s)

    def f64(self):
        return (self.u64() >> 11) * (1.0 / 9007199254740992.0)

    def below(self, n):
        return self.u64() % n


def indexcenter(l, a, b):
    """getcenter(ind)  qtrees.jl:24"""
    if l == 1:
        return a, b
    r, h = 1 << (l - 1), 1 << (l - 2)
    return r * (a - 1) + h, r * (b - 1) + h


class HostGround:
    """a ShiftedQTree on the host: per level (codes[kh, kw], rshift, cshift); reads outside a kernel give `default`"""

    def __init__(self, levels, default):
        self.levels, self.default = levels, int(default)  # levels[l-1] = [codes, rs, cs]
        self.canvas = 1 << (len(levels) - 1)  # the padded side of level 1 (L = log2(canvas) + 1 levels)

    def __len__(self):
        return len(self.levels)

    def size(self, l):
        """size(ground[l], 1): the padded (virtual) side of level l"""
        return self.canvas >> (l - 1)

    @classmethod
    def from_device(cls, d: DeviceQTrees, default):
        levels = []
        for l in range(1, d.nlevels + 1):
            info = np.zeros(5, np.int32)
            d._ck(d.lib.stf_ground_level(d.h, l, None, ptr(info)))
            out = np.zeros((int(info[0]), int(info[1])), np.uint8, order="F")
            d._ck(d.lib.stf_ground_level(d.h, l, ptr(out), ptr(info)))
            levels.append([out, int(info[2]), int(info[3])])
        return cls(levels, default)

    def get(self, l, a, b):
        codes, rs, cs = self.levels[l - 1]
        r, c = a - rs - 1, b - cs - 1
        if 0 <= r < codes.shape[0] and 0 <= c < codes.shape[1]:
            return int(codes[r, c])
        return self.default

    def _rebuild_box(self, l, r0, r1, c0, c1):
        """recompute level l rows [r0, r1) x cols [c0, c1) from level l-1 (qcode of the 4 children, safe reads)"""
        child, crs, ccs = self.levels[l - 2]
        codes, rs, cs = self.levels[l - 1]
        if r1 <= r0 or c1 <= c0:
            return
        er, ec = crs - 2 * rs, ccs - 2 * cs
        rr = 2 * np.arange(r0, r1) - er
        cc = 2 * np.arange(c0, c1) - ec
        acc = np.zeros((r1 - r0, c1 - c0), np.uint8)
        for dr in (0, 1):
            for dc in (0, 1):
                ri, ci = rr + dr, cc + dc
                okr, okc = (ri >= 0) & (ri < child.shape[0]), (ci >= 0) & (ci < child.shape[1])
                v = np.full(acc.shape, self.default, np.uint8)
                sub = child[np.ix_(ri[okr], ci[okc])]
                v[np.ix_(okr, okc)] = sub
                acc |= v
        codes[r0:r1, c0:c1] = acc

    def overlap(self, codes1, rs1, cs1):
        """overlap!(ground, tree): the elementwise rule at level 1 inside the ground kernel, then the touched part of
        every upper level (qtree_functions.jl:452-497)"""
        g, grs, gcs = self.levels[0]
        r0, c0 = rs1 - grs, cs1 - gcs
        a0, a1 = max(0, r0), min(g.shape[0], r0 + codes1.shape[0])
        b0, b1 = max(0, c0), min(g.shape[1], c0 + codes1.shape[1])
        if a1 <= a0 or b1 <= b0:
            return
        t = codes1[a0 - r0:a1 - r0, b0 - c0:b1 - c0]
        cur = g[a0:a1, b0:b1]
        g[a0:a1, b0:b1] = ((cur ^ EMPTY) | (t ^ EMPTY)) ^ EMPTY
        for l in range(2, len(self.levels) + 1):
            codes, rs, cs = self.levels[l - 1]
            _c, crs, ccs = self.levels[l - 2]
            er, ec = crs - 2 * rs, ccs - 2 * cs
            a0, a1 = max(0, (a0 + er - 1) // 2), min(codes.shape[0], (a1 - 1 + er) // 2 + 1)
            b0, b1 = max(0, (b0 + ec - 1) // 2), min(codes.shape[1], (b1 - 1 + ec) // 2 + 1)
            self._rebuild_box(l, a0, a1, b0, b1)


def findroom_uniform(ground: HostGround, q: list, rng: Rng):
    """findroom_uniform(ground, q)  qtree_functions.jl:372-402; q persists between calls (list used as a FIFO)"""
    if not q:
        q.append((len(ground), 1, 1))
    while q:
        i = q.pop(0)
        if i[0] == 1:
            if ground.get(*i) == EMPTY:
                return i
            continue
        ans = None
        for cn in PERM4[rng.below(24)]:
            ci = (i[0] - 1, 2 * i[1] - (cn & 1), 2 * i[2] - (cn >> 1))
            gcode = ground.get(*ci)
            if gcode == EMPTY:
                if rng.f64() < 0.7:  # avoid the local optimum every time
                    ans = ci
                q.append(ci)
            elif gcode == MIX:
                q.append(ci)
        if q and q[0][0] != i[0]:
            for k in range(len(q) - 1, 0, -1):  # shuffle!(q)
                j = rng.u64() % (k + 1)
                q[k], q[j] = q[j], q[k]
        if ans is not None:
            return ans
    return None


def _pnorm_term(x, p):
    if p == 1:
        return x
    if p == 2:
        return x * x
    if p == 3:
        return x * x * x
    if p == 4:
        y = x * x
        return y * y
    return x ** float(p)


def _shufflesort(q: list, ground: HostGround, p, rng: Rng):
    """shufflesort!(q, ground, p)  qtree_functions.jl:403-409: shuffle, then a stable sort by the elliptic p-norm of the
    offset from the level's centre (the padded size of the level, the kernel size of level 1)"""
    ce = (1 + ground.size(q[0][0])) / 2
    h, w = ground.levels[0][0].shape
    for k in range(len(q) - 1, 0, -1):
        j = rng.u64() % (k + 1)
        q[k], q[j] = q[j], q[k]
    q.sort(key=lambda i: _pnorm_term(abs((i[1] - ce) / h), p) + _pnorm_term(abs(i[2] - ce) / w, p))


def findroom_gathering(ground: HostGround, q: list, rng: Rng, level=5, p=2):
    """findroom_gathering(ground, q; level, p)  qtree_functions.jl:410-441: the search starts from every non-FULL cell of
    level max(1, L - level), nearest to the centre first, and re-sorts the queue the same way at every level change"""
    if not q:
        l = max(1, len(ground) - level)
        s = ground.size(l)
        q.extend((l, i, j) for i in range(1, s + 1) for j in range(1, s + 1) if ground.get(l, i, j) != FULL)
        if q:
            _shufflesort(q, ground, p, rng)
    while q:
        i = q.pop(0)
        if i[0] == 1:
            if ground.get(*i) == EMPTY:
                return i
            continue
        ans = None
        for cn in PERM4[rng.below(24)]:
            ci = (i[0] - 1, 2 * i[1] - (cn & 1), 2 * i[2] - (cn >> 1))
            gcode = ground.get(*ci)
            if gcode == EMPTY:
                if rng.f64() < 0.7:
                    ans = ci
                q.append(ci)
            elif gcode == MIX:
                q.append(ci)
        if q and q[0][0] != i[0]:
            _shufflesort(q, ground, p, rng)
        if ans is not None:
            return ans
    return None


def place(d: DeviceQTrees, inds=None, seed=1, mask_label=1, roomfinder="uniform", level=5, p=2):
    """place!(qtrees[, inds])  Stuffing.jl:64-73: ground = deepcopy(mask) + every tree NOT in `inds` (device), then the
    trees of `inds` (default: all but the mask) are centred one after the other on the room the search finds and
    overlapped into the ground.  Moves the trees on the device; returns the list of placed labels."""
    labels = [lb for lb in range(1, len(d) + 1) if lb != mask_label] if inds is None else [int(x) for x in inds]
    ex = i32(labels)
    d._ck(d.lib.stf_ground_compose(d.h, mask_label, len(ex), ptr(ex)))
    ground = HostGround.from_device(d, d.defaults[mask_label - 1])
    rng, q = Rng(seed), []
    new_rs, new_cs = [], []
    for lb in labels:
        ind = findroom_gathering(ground, q, rng, level, p) if roomfinder == "gathering" else findroom_uniform(ground, q, rng)
        if ind is None:
            raise RuntimeError("no room for placement")
        ca, cb = indexcenter(*ind)
        codes1 = d[lb - 1].level(1)
        rs, cs = ca - codes1.shape[0] // 2, cb - codes1.shape[1] // 2  # setcenter!(qtree, getcenter(ind))  qtrees.jl:293-294
        new_rs.append(rs); new_cs.append(cs)
        ground.overlap(codes1, rs, cs)
    if labels:
        d.setshift(labels, 1, new_rs, new_cs)
    return labels


def overlap_ground(d: DeviceQTrees, exclude=(), mask_label=1) -> HostGround:
    """overlap(qtrees) / the ground place! starts from: deepcopy(mask) + every tree not in `exclude`  Stuffing.jl:74-75"""
    ex = i32(list(exclude))
    d._ck(d.lib.stf_ground_compose(d.h, mask_label, len(ex), ptr(ex)))
    return HostGround.from_device(d, d.defaults[mask_label - 1])
