"""ctypes front-end of the CPU ORACLE (oracle/stuffing_oracle.c). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (stuffing.jl_b200/) never does.

The functions mirror the reference's Julia names (qtree, maskqtree, qtrees, collision,
totalcollisions_spacial, partialcollisions, locate!, grad2d, batchsteps!, place!), citing
/root/reference/src/Stuffing.jl and friends.  Labels and (l,a,b) are 1-based.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

EMPTY, FULL, MIX = 1, 2, 3


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("stuffing_oracle.c", "stuffing_oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _LIB_PATH


class Item(C.Structure):
    _fields_ = [("i1", C.c_int32), ("i2", C.c_int32), ("l", C.c_int32), ("a", C.c_int32), ("b", C.c_int32)]


class CellRec(C.Structure):
    _fields_ = [("l", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("label", C.c_int32)]


ITEM_DT = np.dtype([("i1", "<i4"), ("i2", "<i4"), ("l", "<i4"), ("a", "<i4"), ("b", "<i4")])
CELL_DT = np.dtype([("l", "<i4"), ("a", "<i4"), ("b", "<i4"), ("label", "<i4")])

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
        sig = {
            "orc_tree_new": (vp, [vp, i32, i32, i32, i32, i32]),
            "orc_tree_copy": (vp, [vp]),
            "orc_tree_free": (None, [vp]),
            "orc_tree_length": (i32, [vp]),
            "orc_tree_sizelevel": (i32, [vp]),
            "orc_tree_default": (i32, [vp]),
            "orc_tree_level_info": (None, [vp, i32, vp]),
            "orc_tree_level_data": (None, [vp, i32, vp]),
            "orc_tree_get": (i32, [vp, i32, i32, i32]),
            "orc_tree_set": (i32, [vp, i32, i32, i32, i32]),
            "orc_level_setshift": (None, [vp, i32, i32, i32]),
            "orc_build": (None, [vp, i32]),
            "orc_shift": (None, [vp, i32, i32, i32]),
            "orc_setshift": (None, [vp, i32, i32, i32]),
            "orc_shift_axis": (None, [vp, i32, i32, i32, i32]),
            "orc_setshifts": (i64, [vp, vp, i32, i32, vp, vp, i32, i32]),
            "orc_getshifts": (None, [vp, vp, i32, i32, vp, vp]),
            "orc_testqtree": (i32, [vp]),
            "orc_inkernelbounds": (i32, [vp, i32, i32, i32]),
            "orc_collision_dfs": (None, [vp, vp, i32, i32, i32, vp]),
            "orc_collision_bfs": (None, [vp, vp, i32, i32, i32, u64, vp, vp]),
            "orc_collision": (None, [vp, vp, u64, vp]),
            "orc_collide_items": (i64, [vp, vp, i64, u64, i32, vp, i64, vp]),
            "orc_totalcollisions_native": (i64, [vp, i32, vp, i32, u64, vp, i64]),
            "orc_locate_hash": (i64, [vp, i32, vp, i32, vp, i64]),
            "orc_totalcollisions_spacial": (i64, [vp, i32, vp, i32, i32, u64, i32, vp, i64, vp, i64, vp, vp, vp]),
            "orc_sort_unique": (None, [vp, vp, i32]),
            "orc_linked_new": (vp, [i32]),
            "orc_linked_free": (None, [vp]),
            "orc_linked_locate": (None, [vp, vp, vp, i32]),
            "orc_linked_clear": (None, [vp, i32]),
            "orc_linked_collect": (i64, [vp, vp, i64]),
            "orc_partialcollisions": (i64, [vp, i32, vp, vp, i32, i32, u64, i32, vp, i64, vp, i64, vp, vp]),
            "orc_grad2d": (None, [vp, i32, i32, i32, vp]),
            "orc_intlog2": (i32, [C.c_double]),
            "orc_opt_new": (vp, [i32, C.c_double, C.c_double, i32]),
            "orc_opt_free": (None, [vp]),
            "orc_opt_set_eta": (None, [vp, C.c_double]),
            "orc_opt_reset": (None, [vp, i32]),
            "orc_opt_velocity": (None, [vp, i32, vp, vp]),
            "orc_batchsteps": (None, [vp, i32, vp, i64, vp, u64, u64]),
            "orc_batchsteps_masks": (None, [vp, i32, vp, i64, vp, u64, u64, vp]),
            "orc_place": (i32, [vp, i32, u64]),
            "orc_place_gathering": (i32, [vp, i32, u64, i32, C.c_double]),
            "orc_overlap": (None, [vp, vp]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OTree:
    """ShiftedQTree / U8SQTree (qtrees.jl:170-294)."""

    def __init__(self, pic: np.ndarray, canvas: int, default: int = EMPTY, _h=None):
        if _h is not None:
            self.h = _h
            return
        pic = np.asfortranarray(pic, dtype=np.uint8)
        kh, kw = pic.shape
        self.h = lib().orc_tree_new(_ptr(pic), kh, kw, max(kh, 1), int(canvas), int(default))

    def __del__(self):
        try:
            lib().orc_tree_free(self.h)
        except Exception:
            pass

    def copy(self):
        return OTree(None, 0, _h=lib().orc_tree_copy(self.h))

    def __len__(self):
        return lib().orc_tree_length(self.h)

    @property
    def default(self):
        return lib().orc_tree_default(self.h)

    def info(self, l=1):
        out = np.zeros(5, np.int32)
        lib().orc_tree_level_info(self.h, l, _ptr(out))
        return tuple(int(x) for x in out)  # kh, kw, rs, cs, canvas

    def level(self, l):
        kh, kw, *_ = self.info(l)
        out = np.zeros((kh, kw), np.uint8, order="F")
        if kh * kw:
            lib().orc_tree_level_data(self.h, l, _ptr(out))
        return out

    def get(self, l, a, b):
        return lib().orc_tree_get(self.h, l, a, b)

    def set(self, l, a, b, v):
        if lib().orc_tree_set(self.h, l, a, b, v) != 0:
            raise IndexError("BoundsError")

    def getshift(self, l=1):
        return self.info(l)[2:4]

    def kernelsize(self, l=1):
        return self.info(l)[0:2]

    def getcenter(self):
        (rs, cs), (kh, kw) = self.getshift(), self.kernelsize()
        return rs + kh // 2, cs + kw // 2

    def setcenter(self, c):
        kh, kw = self.kernelsize()
        self.setshift(1, c[0] - kh // 2, c[1] - kw // 2)

    def build(self, layer=2):
        lib().orc_build(self.h, layer)
        return self

    def shift(self, l, s1, s2):
        lib().orc_shift(self.h, l, s1, s2)

    def setshift(self, l, s1, s2):
        lib().orc_setshift(self.h, l, s1, s2)

    def level_setshift(self, l, rs, cs):
        lib().orc_level_setshift(self.h, l, rs, cs)

    def testqtree(self):
        return lib().orc_testqtree(self.h)

    def inkernelbounds(self, l, a, b):
        return bool(lib().orc_inkernelbounds(self.h, l, a, b))


def _canvas_for(shape):
    m = max(2, *shape)  # qtrees.jl:201-204
    return 2 ** math.ceil(math.log2(m))


def qtree(pic, size=None, background=None):
    """qtree (Stuffing.jl:16-27): uint8 → codes as given; bool → FULL/EMPTY; other → != background."""
    pic = np.asarray(pic)
    if pic.dtype == np.uint8:
        codes = pic
    elif pic.dtype == np.bool_:
        codes = np.where(pic, FULL, EMPTY).astype(np.uint8)
    else:
        bg = pic.flat[0] if background is None else background
        codes = np.where(pic != bg, FULL, EMPTY).astype(np.uint8)
    canvas = _canvas_for(codes.shape) if size is None else size
    return OTree(codes, canvas, EMPTY).build()


def maskqtree(pic, background=None):
    """maskqtree (Stuffing.jl:29-46)."""
    pic = np.asarray(pic)
    if pic.dtype == np.uint8:
        codes = pic
    elif pic.dtype == np.bool_:
        codes = np.where(pic, EMPTY, FULL).astype(np.uint8)
    else:
        bg = pic.flat[0] if background is None else background
        codes = np.where(pic != bg, EMPTY, FULL).astype(np.uint8)
    ms = max(codes.shape)
    b = max(ms * 0.024, 20)
    s = 2 ** math.ceil(math.log2(ms + b))
    t = OTree(codes, s, FULL)
    t.level_setshift(1, (s - codes.shape[0]) // 2, (s - codes.shape[1]) // 2)
    return t.build()


def qtrees(pics, mask=None, size=None, background=None, maskbackground=None):
    """qtrees (Stuffing.jl:48-62); the first tree is the mask when one is given."""
    ts = []
    if mask is not None:
        mq = maskqtree(mask, background=maskbackground)
        ts.append(mq)
        if size is None:
            size = mq.info(1)[4]
    elif size is None:
        size = 2 ** math.ceil(math.log2(max(max(np.shape(p)) for p in pics)))
    for p in pics:
        ts.append(qtree(p, size, background=background))
    return ts


class TreeList(list):
    """a list of OTrees that keeps its ctypes handle array (timed loops: no per-call marshalling of 10k handles);
    do not mutate it after the first call"""
    _harr = None


def _handles(qts):
    if isinstance(qts, TreeList):
        if qts._harr is None or len(qts._harr) != max(1, len(qts)):
            qts._harr = (C.c_void_p * max(1, len(qts)))(*[t.h for t in qts])
        return qts._harr
    arr = (C.c_void_p * max(1, len(qts)))(*[t.h for t in qts])
    return arr


def setshifts(qts, labels, level, rs, cs, skip_same=True, nthreads=1):
    """setshift!(qts[labels[i]], level, (rs[i], cs[i])) for the whole list in one C call; returns the number of trees rebuilt"""
    lb, nl = _labels(labels)
    rs = np.ascontiguousarray(rs, np.int32); cs = np.ascontiguousarray(cs, np.int32)
    if lb is None:
        nl = len(rs)
    return int(lib().orc_setshifts(_handles(qts), _ptr(lb), nl, level, _ptr(rs), _ptr(cs), int(skip_same), nthreads))


def getshifts(qts, labels=None, level=1):
    lb, nl = _labels(labels)
    if lb is None:
        nl = len(qts)
    rs = np.zeros(nl, np.int32); cs = np.zeros(nl, np.int32)
    lib().orc_getshifts(_handles(qts), _ptr(lb), nl, level, _ptr(rs), _ptr(cs))
    return rs, cs


def _labels(labels):
    if labels is None:
        return None, 0
    if isinstance(labels, np.ndarray) and labels.dtype == np.int32 and labels.flags.c_contiguous:
        return labels, len(labels)
    a = np.ascontiguousarray(np.asarray(list(labels), dtype=np.int32))
    return a, len(a)


def _items_out(arr, n):
    return [((int(r["i1"]), int(r["i2"])), (int(r["l"]), int(r["a"]), int(r["b"]))) for r in arr[:n]]


def items_array(items):
    a = np.zeros(len(items), ITEM_DT)
    for k, ((i1, i2), (l, x, y)) in enumerate(items):
        a[k] = (i1, i2, l, x, y)
    return a


def collision_dfs(q1, q2, at=None):
    at = at or (len(q1), 1, 1)
    out = np.zeros(3, np.int32)
    lib().orc_collision_dfs(q1.h, q2.h, at[0], at[1], at[2], _ptr(out))
    return tuple(int(x) for x in out)


def collision_bfs(q1, q2, at=None, perm_seed=0, return_visited=False):
    at = at or (len(q1), 1, 1)
    out = np.zeros(3, np.int32)
    vis = C.c_int64(0)
    lib().orc_collision_bfs(q1.h, q2.h, at[0], at[1], at[2], perm_seed, _ptr(out), C.byref(vis))
    r = tuple(int(x) for x in out)
    return (r, vis.value) if return_visited else r


def collision(q1, q2, perm_seed=0):
    assert len(q1) == len(q2)
    out = np.zeros(3, np.int32)
    lib().orc_collision(q1.h, q2.h, perm_seed, _ptr(out))
    return tuple(int(x) for x in out)


def collide_items(qts, items, perm_seed=0, nthreads=1, raw=False):
    arr = items if isinstance(items, np.ndarray) else items_array(items)
    out = np.zeros(max(1, len(arr)), ITEM_DT)
    vis = C.c_int64(0)
    n = lib().orc_collide_items(_handles(qts), _ptr(arr), len(arr), perm_seed, nthreads, _ptr(out), len(out), C.byref(vis))
    if raw:
        return out[:n], vis.value
    return _items_out(out, n)


def totalcollisions_native(qts, labels=None, perm_seed=0):
    lb, nl = _labels(labels)
    n = len(qts)
    cap = max(16, n * (n - 1) // 2 if labels is None else nl * nl)
    out = np.zeros(cap, ITEM_DT)
    m = lib().orc_totalcollisions_native(_handles(qts), n, _ptr(lb), nl, perm_seed, _ptr(out), cap)
    assert m >= 0
    return _items_out(out, m)


def locate_hash(qts, labels=None):
    """locate!(qts[, labels], HashSpacialQTree) → {(l,a,b): [labels in insertion order]}"""
    lb, nl = _labels(labels)
    cap = 4 * len(qts) + 4
    out = np.zeros(cap, CELL_DT)
    m = lib().orc_locate_hash(_handles(qts), len(qts), _ptr(lb), nl, _ptr(out), cap)
    d = {}
    for r in out[:m]:
        d.setdefault((int(r["l"]), int(r["a"]), int(r["b"])), []).append(int(r["label"]))
    return d


def totalcollisions_spacial(qts, labels=None, unique=True, perm_seed=0, nthreads=1, details=False, raw=False):
    lb, nl = _labels(labels)
    n = len(qts)
    cap = 1 << 12
    while True:
        out = np.zeros(cap, ITEM_DT)
        items = np.zeros(cap if details else 1, ITEM_DT)
        ni, nd, vis = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        m = lib().orc_totalcollisions_spacial(
            _handles(qts), n, _ptr(lb), nl, int(unique), perm_seed, nthreads, _ptr(out), cap,
            _ptr(items) if details else None, cap if details else 0, C.byref(ni), C.byref(nd), C.byref(vis))
        if m >= 0 and (not details or ni.value <= cap):
            break
        cap = max(-m if m < 0 else 0, ni.value, cap * 2) + 16
    if raw:
        return out[:m], dict(n_items=ni.value, n_direct=nd.value, visited=vis.value, items=items[: ni.value] if details else None)
    res = _items_out(out, m)
    if details:
        return res, dict(items=_items_out(items, ni.value), n_direct=nd.value, visited=vis.value)
    return res


class LinkedTree:
    """LinkedSpacialQTree (qtrees.jl:424-552) — linked_spacial_qtree(qts)."""

    def __init__(self, qts):
        self.h = lib().orc_linked_new(len(qts[0]) if len(qts) else 0)

    def __del__(self):
        try:
            lib().orc_linked_free(self.h)
        except Exception:
            pass

    def locate(self, qts, labels=None):
        lb, nl = _labels(range(1, len(qts) + 1) if labels is None else labels)
        lib().orc_linked_locate(self.h, _handles(qts), _ptr(lb), nl)
        return self

    def clear(self, label):
        lib().orc_linked_clear(self.h, label)

    def collect(self):
        cap = 1 << 12
        while True:
            out = np.zeros(cap, CELL_DT)
            m = lib().orc_linked_collect(self.h, _ptr(out), cap)
            if m >= 0:
                break
            cap = -m
        d = {}
        for r in out[:m]:
            d.setdefault((int(r["l"]), int(r["a"]), int(r["b"])), []).append(int(r["label"]))
        return d


def partialcollisions(qts, tree=None, labels=None, unique=True, perm_seed=0, nthreads=1, details=False, raw=False):
    tree = tree if tree is not None else LinkedTree(qts)
    lb, nl = _labels(range(1, len(qts) + 1) if labels is None else labels)
    cap = 1 << 12
    while True:
        out = np.zeros(cap, ITEM_DT)
        items = np.zeros(cap if details else 1, ITEM_DT)
        ni, nd = C.c_int64(0), C.c_int64(0)
        m = lib().orc_partialcollisions(
            _handles(qts), len(qts), tree.h, _ptr(lb), nl, int(unique), perm_seed, nthreads, _ptr(out), cap,
            _ptr(items) if details else None, cap if details else 0, C.byref(ni), C.byref(nd))
        if m >= 0 and (not details or ni.value <= cap):
            break
        cap = max(-m if m < 0 else 0, ni.value, cap * 2) + 16
    if raw:
        return out[:m], dict(n_items=ni.value, n_direct=nd.value)
    res = _items_out(out, m)
    if details:
        return res, dict(items=_items_out(items, ni.value), n_direct=nd.value)
    return res


SPACIAL_ENABLE_THRESHOLD = 10  # round(Int, 10 + 10log2(nthreads)) with one thread  qtree_functions.jl:265


def totalcollisions(qts, labels=None, **kw):
    """dispatcher  qtree_functions.jl:270-276"""
    last = qts if labels is None else labels
    if len(last) > SPACIAL_ENABLE_THRESHOLD:
        return totalcollisions_spacial(qts, labels, **kw)
    kw.pop("unique", None)
    return totalcollisions_native(qts, labels, **kw)


def dynamiccollisions(qts, tree, moved: list, unlocated: list, **kw):
    """dynamiccollisions  qtree_functions.jl:340-356; `moved`/`unlocated` are ordered sets (lists, in place)."""
    if len(moved) / len(qts) > 0.6:
        r = totalcollisions(qts, **kw)
        for m in moved:
            if m not in unlocated:
                unlocated.append(m)
    else:
        ms = set(moved)
        tree.locate(qts, [i for i in unlocated if i not in ms])
        unlocated.clear()
        r = partialcollisions(qts, tree, moved, **kw)
    moved.clear()
    for (i1, i2), _ in r:
        for x in (i1, i2):
            if x not in moved:
                moved.append(x)
    return r


def grad2d(t, l, a, b):
    out = np.zeros(2, np.int32)
    lib().orc_grad2d(t.h, l, a, b, _ptr(out))
    return int(out[0]), int(out[1])


def intlog2(x):
    return lib().orc_intlog2(float(x))


class Opt:
    """kind: 'default' (Δ/6, fit.jl:25), 'sgd', 'momentum' (fit_helper.jl:13-38)."""

    KIND = {"default": 0, "sgd": 1, "momentum": 2}

    def __init__(self, kind="momentum", eta=0.25, rho=0.5, n=0):
        self.h = lib().orc_opt_new(self.KIND[kind], eta, rho, n)

    def __del__(self):
        try:
            lib().orc_opt_free(self.h)
        except Exception:
            pass

    def set_eta(self, eta):
        lib().orc_opt_set_eta(self.h, eta)

    def reset(self, label):
        lib().orc_opt_reset(self.h, label)

    def velocity(self, label):
        out = np.zeros(2, np.float64)
        has = C.c_int32(0)
        lib().orc_opt_velocity(self.h, label, _ptr(out), C.byref(has))
        return (float(out[0]), float(out[1])), bool(has.value)


def batchsteps(qts, colist, opt: Opt | None, seed=0, iteration=0, is_mask=None):
    arr = colist if isinstance(colist, np.ndarray) else items_array(colist)
    im = None if is_mask is None else np.ascontiguousarray(is_mask, np.uint8)
    lib().orc_batchsteps_masks(_handles(qts), len(qts), _ptr(arr), len(arr), opt.h if opt else None, seed, iteration, _ptr(im))


def place(qts, seed=1, roomfinder="uniform", level=5, p=2):
    if roomfinder == "gathering":
        r = lib().orc_place_gathering(_handles(qts), len(qts), seed, level, float(p))
    else:
        r = lib().orc_place(_handles(qts), len(qts), seed)
    if r != 0:
        raise RuntimeError("no room for placement")
    return qts


def overlap(ground, t):
    lib().orc_overlap(ground.h, t.h)
