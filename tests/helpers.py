"""Shared helpers of the parity tests: build the same scene for the CPU oracle and for the device."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import _pkg  # noqa: E402
import workloads  # noqa: E402


def product():
    return _pkg.load()


def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


def device_from_oracle(S, qts_orc, **kw):
    """a DeviceQTrees with the same level-1 payloads, defaults and level-1 shifts as the oracle trees.
    Only valid while every oracle tree is in its freshly built state (shift chain from level 1)."""
    hts = []
    for t in qts_orc:
        kh, kw_, rs, cs, canvas = t.info(1)
        hts.append(S.HostTree(t.level(1), canvas, t.default, (rs, cs)))
    return S.DeviceQTrees(hts, **kw)


def assert_same_trees(d, qts_orc, labels=None):
    """every level of every listed tree: same kernel size, same shift, same node codes.  The device side is read with one
    batched download (stf_download_trees) and one stf_get_shifts per level — a handful of launches, not one per level"""
    labels = list(range(1, len(qts_orc) + 1)) if labels is None else list(labels)
    if not labels:
        return
    L = d.nlevels
    dev = d.download_trees(labels)
    shifts = [d.getshifts(labels, l) for l in range(1, L + 1)]
    for k, lb in enumerate(labels):
        to = qts_orc[lb - 1]
        assert len(to) == L
        for l in range(1, L + 1):
            io = to.info(l)
            ld = dev[k][l - 1]
            idv = (ld.shape[0], ld.shape[1], int(shifts[l - 1][0][k]), int(shifts[l - 1][1][k]), d.canvas >> (l - 1))
            assert tuple(io) == idv, (lb, l, io, idv)
            assert np.array_equal(to.level(l), ld), (lb, l)


def rect_scene(orc, n, seed, sz=100, masksize=(1000, 1000), place=True):
    """randqtrees(n; sz, masksize) of src/utils.jl:35-39 on the oracle"""
    rng = np.random.default_rng(seed)
    objs = workloads.rect_objects(n, sz, rng)
    qts = orc.qtrees(objs, mask=np.ones(masksize, bool))
    if place:
        orc.place(qts, seed)
    return qts


def pairset(cl):
    return {frozenset(p) for p, _ in cl}
