"""Wire formats at the boundary of the hot path.

  * `charmat` / `charimage`  — the reference's text rendering of a level (src/qtrees.jl:308-345), glyph for glyph:
    EMPTY "░", FULL "▓", MIX "▒", anything else "▞", long axes elided around the middle;
  * `save_scene` / `load_scene` — a directory of plain `.npy` files (NumPy format 1.0, little endian, Fortran order for
    matrices = Julia's column-major memory) holding the level-1 payloads, defaults and level-1 shifts of a vector of
    trees.  julia/NpyIO.jl reads and writes the same files, so a scene fitted on the GPU can be checked against a real
    Julia run of the reference on another box, and vice versa;
  * `packing` — packing(mask, objs, ...) of src/Stuffing.jl:101-117: qtrees -> place! -> fit! -> getpositions.
"""
from __future__ import annotations

import json
import math
import os

import numpy as np

from .qtrees import EMPTY, FULL, MIX, DeviceQTrees, DeviceTree, HostTree, getpositions, qtrees as _qtrees

_GLYPH = {EMPTY: "░", FULL: "▓", MIX: "▒"}


def charmat(mat: np.ndarray, maxlen: int = 49):
    """charmat(mat; maxlen)  qtrees.jl:309-336 — a list of rows (lists of one-character strings)"""
    mat = np.asarray(mat)
    m, n = mat.shape
    half = maxlen // 2
    maxlen = half * 2 + 1
    if m <= maxlen and n <= maxlen:
        return [[_GLYPH.get(int(x), "▞") for x in row] for row in mat]
    if m > maxlen:
        _n = min(maxlen, n) - 5
        if _n > 0:
            _n2 = _n - _n // 2
            vsp = ["⋮"] + [" "] * (_n2 // 2) + ["⋮"] + [" "] * (_n2 - _n2 // 2) + ["⋱"] + [" "] * (_n // 4) + ["⋮"] + [" "] * (_n // 2 - _n // 4) + ["⋮"]
        else:
            vsp = ["⋮"] * min(maxlen, n)
        return charmat(mat[:half, :], maxlen) + [vsp] + charmat(mat[m - half:, :], maxlen)
    left, right = charmat(mat[:, :half], maxlen), charmat(mat[:, n - half:], maxlen)
    return [l + ["…"] + r for l, r in zip(left, right)]


def charimage(x, maxlen: int = 49) -> str:
    """charimage(mat) / charimage(qt::ShiftedQTree)  qtrees.jl:338-339: a tree is shown at the first level that fits"""
    if isinstance(x, (DeviceTree, HostTree)):
        side = x.codes.shape[0] if isinstance(x, HostTree) else x.info(1)[4]
        level = max(1, math.ceil(math.log2(side / maxlen)) + 1)
        if isinstance(x, HostTree):
            if level != 1:
                raise ValueError("a HostTree only holds level 1 before upload")
            mat = x.codes
        else:  # the canvas view of the level: default outside the kernel  (PaddedMat getindex, qtrees.jl:156-162)
            kh, kw, rs, cs, s = x.info(level)
            mat = np.full((s, s), x.default, np.uint8)
            k = x.level(level)
            r0, r1, c0, c1 = max(0, rs), min(s, rs + kh), max(0, cs), min(s, cs + kw)
            if r1 > r0 and c1 > c0:
                mat[r0:r1, c0:c1] = k[r0 - rs:r1 - rs, c0 - cs:c1 - cs]
        return charimage(mat, maxlen)
    return "".join("".join(row) + "\n" for row in charmat(x, maxlen))


# ------------------------------------------------------------------------------------------------ .npy scene exchange
def save_scene(path: str, qts, positions_only: bool = False) -> None:
    """a vector of trees -> `path`/: meta.json, shifts.npy (n x 2 int32: rshift, cshift of level 1), defaults.npy (n uint8),
    tree_00001.npy ... (kh x kw uint8 level-1 codes, Fortran order).  Works on a DeviceQTrees (state read from the GPU) or on
    a list of HostTrees."""
    os.makedirs(path, exist_ok=True)
    if isinstance(qts, DeviceQTrees):
        rs, cs = qts.getshifts()
        shifts = np.stack([rs, cs], 1).astype("<i4")
        defaults = np.asarray(qts.defaults, np.uint8)
        canvas, n = qts.canvas, len(qts)
        codes = None if positions_only else [lv[0] for lv in qts.download_trees()]
    else:
        qts = list(qts)
        shifts = np.array([t.shift1 for t in qts], "<i4").reshape(-1, 2)
        defaults = np.array([t.default for t in qts], np.uint8)
        canvas, n = (qts[0].canvas if qts else 2), len(qts)
        codes = None if positions_only else [t.codes for t in qts]
    np.save(os.path.join(path, "shifts.npy"), shifts)
    np.save(os.path.join(path, "defaults.npy"), defaults)
    if codes is not None:
        for i, c in enumerate(codes):
            np.save(os.path.join(path, f"tree_{i + 1:05d}.npy"), np.asfortranarray(c.astype(np.uint8)))
    with open(os.path.join(path, "meta.json"), "w") as f:
        json.dump({"format": "stuffing-scene-npy-1", "n": n, "canvas": int(canvas), "codes": {"EMPTY": EMPTY, "FULL": FULL, "MIX": MIX},
                   "order": "matrices are Fortran (column-major) order, labels 1-based, label 1 = the mask by convention"}, f)


def load_scene(path: str):
    """`path`/ -> list of HostTrees at the stored level-1 shifts (upload with DeviceQTrees(trees))"""
    meta = json.load(open(os.path.join(path, "meta.json")))
    shifts = np.load(os.path.join(path, "shifts.npy"))
    defaults = np.load(os.path.join(path, "defaults.npy"))
    trees = []
    for i in range(meta["n"]):
        c = np.load(os.path.join(path, f"tree_{i + 1:05d}.npy"))
        trees.append(HostTree(np.asfortranarray(c), meta["canvas"], int(defaults[i]), (int(shifts[i, 0]), int(shifts[i, 1]))))
    return trees


# ------------------------------------------------------------------------------------------------ packing (Stuffing.jl:101-117)
def packing(mask, objs, epochs=-1, background=None, maskbackground=None, device=0, seed=1, **kargs):
    """packing(mask, objs, args...; kargs...): qtrees -> place! -> fit! -> getpositions; returns the positions"""
    from .place import place
    from .trainer import train
    d = _qtrees(objs, mask=mask, background=background, maskbackground=maskbackground, device=device)
    place(d, seed=seed)
    ep, nc = train(d, epochs, seed=seed, **kargs)
    return getpositions(d), (ep, nc)
