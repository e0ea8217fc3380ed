"""Host-side mirror of the reference's QTrees interface over the C ABI.

Names, argument meaning and error behaviour follow /root/reference/src/qtrees.jl,
qtree_functions.jl and Stuffing.jl (Julia's trailing `!` is dropped).  A Python list of trees is
0-based, but every label and (l,a,b) that crosses the boundary is 1-based, exactly as the
reference prints them: `((i1, i2), (l, a, b))` is a CoItem.

Construction is two-stage because the device owns whole vectors of trees:
  qtree()/maskqtree() return a HostTree (level-1 codes + shift, nothing built);
  qtrees() / DeviceQTrees(list) upload a vector and build every pyramid on the GPU.
Functions that take `qtrees` accept either; a plain list of HostTrees is uploaded on the fly.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import CELL_DT, ITEM_DT, StuffingError, i32, ptr

EMPTY, FULL, MIX = 1, 2, 3
SPACIAL_ENABLE_THRESHOLD = 10  # round(Int, 10 + 10log2(Threads.nthreads())) for one thread  qtree_functions.jl:265


# ------------------------------------------------------------------------------------------------
class HostTree:
    """Arguments of ShiftedQTree(pic, sz; default) before upload (qtrees.jl:196-209)."""

    def __init__(self, codes: np.ndarray, canvas: int, default: int = EMPTY, shift=(0, 0)):
        codes = np.asfortranarray(codes, dtype=np.uint8)
        assert codes.ndim == 2
        assert canvas >= 2 and canvas & (canvas - 1) == 0, "canvas must be a power of two"  # qtrees.jl:197
        self.codes, self.canvas, self.default, self.shift1 = codes, int(canvas), int(default), (int(shift[0]), int(shift[1]))

    def __len__(self):
        return int(math.log2(self.canvas)) + 1

    def setshift(self, *args):  # setshift!(t, [l,] st1, st2) / setshift!(t, [l,] (st1, st2)) before upload: level 1 only
        if len(args) == 1:
            l, (s1, s2) = 1, args[0]
        elif len(args) == 2:
            l, (s1, s2) = args
        else:
            l, s1, s2 = args
        assert l == 1, "a HostTree only carries its level-1 shift; upload first"
        self.shift1 = (int(s1), int(s2))

    def kernelsize(self, l=1):
        assert l == 1
        return self.codes.shape

    def getshift(self, l=1):
        assert l == 1
        return self.shift1


def _codes(pic, true_code, false_code, background):
    pic = np.asarray(pic)
    if pic.dtype == np.uint8:
        return pic  # already node codes (Stuffing.jl:16-19)
    if pic.dtype == np.bool_:
        return np.where(pic, true_code, false_code).astype(np.uint8)
    assert pic.size > 0 or background is not None
    bg = pic.flat[0] if background is None else background
    return np.where(pic != bg, true_code, false_code).astype(np.uint8)


def qtree(pic, size=None, background=None) -> HostTree:
    """qtree(pic[, size]; background)  Stuffing.jl:16-27"""
    codes = _codes(pic, FULL, EMPTY, background)
    if size is None:
        m = max(2, *codes.shape)  # qtrees.jl:201-204
        size = 2 ** math.ceil(math.log2(m))
    return HostTree(codes, size, EMPTY)


def maskqtree(pic, background=None) -> HostTree:
    """maskqtree(pic; background)  Stuffing.jl:29-46: inside = EMPTY, outside and beyond = FULL, centred"""
    codes = _codes(pic, EMPTY, FULL, background)
    ms = max(codes.shape)
    b = max(ms * 0.024, 20)
    s = 2 ** math.ceil(math.log2(ms + b))
    return HostTree(codes, s, FULL, ((s - codes.shape[0]) // 2, (s - codes.shape[1]) // 2))


class DeviceTree:
    """One ShiftedQTree living in a DeviceQTrees (qtrees.jl:170-294)."""

    def __init__(self, owner: "DeviceQTrees", label: int):
        self.owner, self.label = owner, label

    def __len__(self):
        return self.owner.nlevels

    @property
    def default(self):
        return self.owner.defaults[self.label - 1]

    def info(self, l=1):
        out = np.zeros(5, np.int32)
        self.owner._ck(self.owner.lib.stf_get_level_info(self.owner.h, self.label, l, ptr(out)))
        return tuple(int(x) for x in out)

    def kernelsize(self, l=1):
        return self.info(l)[0:2]

    def getshift(self, l=1):
        return self.info(l)[2:4]

    def getcenter(self):  # qtrees.jl:290
        kh, kw, rs, cs, _ = self.info(1)
        return rs + kh // 2, cs + kw // 2

    def level(self, l):
        """t[l].kernel payload as a (kh, kw) uint8 array"""
        kh, kw = self.kernelsize(l)
        out = np.zeros((kh, kw), np.uint8, order="F")
        if kh * kw:
            self.owner._ck(self.owner.lib.stf_download_level(self.owner.h, self.label, l, ptr(out)))
        return out

    def get(self, l, a, b):
        """t[l, a, b]"""
        return int(self.owner.read_nodes([self.label], [l], [a], [b])[0])

    def set(self, l, a, b, v):
        kh, kw, rs, cs, _ = self.info(l)
        if not (rs < a <= rs + kh and cs < b <= cs + kw):
            raise IndexError("BoundsError")  # qtrees.jl:164-166
        raise NotImplementedError("node writes are not part of the hot path")

    def shift(self, l, s1, s2):  # shift!(t, l, st1, st2)
        self.owner.shift([self.label], l, [s1], [s2])

    def setshift(self, *args):  # setshift!(t, l, st1, st2) / setshift!(t, (st1, st2))
        if len(args) == 1:
            l, (s1, s2) = 1, args[0]
        elif len(args) == 2:
            l, (s1, s2) = args
        else:
            l, s1, s2 = args
        self.owner.setshift([self.label], l, [s1], [s2])

    def setcenter(self, c):  # qtrees.jl:293-294
        kh, kw = self.kernelsize()
        self.setshift(1, c[0] - kh // 2, c[1] - kw // 2)

    def build(self, layer=2):  # buildqtree!(t, layer)
        self.owner.build([self.label], layer)
        return self

    def inkernelbounds(self, l, a, b):
        kh, kw, rs, cs, _ = self.info(l)
        return rs < a <= rs + kh and cs < b <= cs + kw


class DeviceQTrees:
    """Vector{U8SQTree} resident on one GPU: a context of the C ABI."""

    def __init__(self, trees, device: int = 0, is_mask=None, scene=None, handle=None):
        """`handle`: upload into an existing context (one of a GpuGroup's) instead of creating one; the group owns it"""
        self.lib = _lib.load()
        trees = list(trees)
        self._owns = handle is None
        if handle is not None:
            self.h = C.c_void_p(handle)
        else:
            self.h = C.c_void_p()
            rc = self.lib.stf_create(device, C.byref(self.h))
            if rc != 0:
                raise StuffingError(rc, "stf_create failed: no CUDA device (there is no CPU fallback)")
        self.n = len(trees)
        self.device = device
        self.defaults = [t.default for t in trees]
        canvases = {t.canvas for t in trees}
        assert len(canvases) <= 1, "all trees of a vector share one canvas"
        self.canvas = canvases.pop() if canvases else 2
        pics = (C.c_void_p * max(1, self.n))(*[t.codes.ctypes.data for t in trees])
        self._keep = trees
        kh, kw = i32([t.codes.shape[0] for t in trees]), i32([t.codes.shape[1] for t in trees])
        self.kh1, self.kw1 = kh, kw
        pitch = i32([max(1, t.codes.shape[0]) for t in trees])
        rs, cs = i32([t.shift1[0] for t in trees]), i32([t.shift1[1] for t in trees])
        dflt = np.asarray(self.defaults, np.uint8)
        im = None if is_mask is None else np.ascontiguousarray(is_mask, np.uint8)
        sc = None if scene is None else i32(scene)
        self._ck(self.lib.stf_upload_objects(self.h, self.n, pics, ptr(kh), ptr(kw), ptr(pitch), ptr(rs), ptr(cs), ptr(dflt),
                                             self.canvas, ptr(im), ptr(sc)))
        self._keep = None
        self.nlevels = self.lib.stf_num_levels(self.h)
        self._cap = 1 << 14
        self._buf = np.zeros(self._cap, ITEM_DT)

    @classmethod
    def from_packed(cls, payload, kh, kw, rs, cs, dflt, canvas, is_mask=None, scene=None, device: int = 0, payload_dev_ptr=None):
        """many trees from ONE concatenated column-major payload (stf_upload_packed): the bulk-upload route for
        batches of scenes.  `payload_dev_ptr`: the payload already lives in HBM (needs 32 bytes of zeroed slack)."""
        self = cls.__new__(cls)
        self.lib = _lib.load()
        self._owns = True
        self.h = C.c_void_p()
        rc = self.lib.stf_create(device, C.byref(self.h))
        if rc != 0:
            raise StuffingError(rc, "stf_create failed: no CUDA device (there is no CPU fallback)")
        kh, kw, rs, cs = i32(kh), i32(kw), i32(rs), i32(cs)
        dflt = np.ascontiguousarray(dflt, np.uint8)
        self.n, self.device, self.canvas, self.defaults = len(kh), device, int(canvas), [int(x) for x in dflt]
        self.kh1, self.kw1 = kh, kw
        im = None if is_mask is None else np.ascontiguousarray(is_mask, np.uint8)
        sc = None if scene is None else i32(scene)
        if payload_dev_ptr is not None:
            src, on_dev = C.c_void_p(payload_dev_ptr), 1
        else:
            payload = np.ascontiguousarray(payload, np.uint8)
            src, on_dev = C.c_void_p(payload.ctypes.data), 0
        self._ck(self.lib.stf_upload_packed(self.h, self.n, src, on_dev, ptr(kh), ptr(kw), ptr(rs), ptr(cs), ptr(dflt), self.canvas, ptr(im), ptr(sc)))
        self._keep = None
        self.nlevels = self.lib.stf_num_levels(self.h)
        self._cap = 1 << 14
        self._buf = np.zeros(self._cap, ITEM_DT)
        return self

    def __del__(self):
        try:
            if self.h and getattr(self, "_owns", True):
                self.lib.stf_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise StuffingError(rc, self.lib.stf_last_error(self.h).decode())

    # list protocol
    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [DeviceTree(self, k + 1) for k in range(*i.indices(self.n))]
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        return DeviceTree(self, i + 1)

    def __iter__(self):
        return (DeviceTree(self, k + 1) for k in range(self.n))

    # state
    def read_nodes(self, labels, l, a, b):
        labels, l, a, b = i32(labels), i32(l), i32(a), i32(b)
        out = np.zeros(len(labels), np.uint8)
        self._ck(self.lib.stf_read_nodes(self.h, len(labels), ptr(labels), ptr(l), ptr(a), ptr(b), ptr(out)))
        return out

    def download_trees(self, labels=None):
        """every level of the listed trees with ONE kernel launch (stf_download_trees): a list (per tree) of lists
        (per level 1..L) of kh x kw uint8 code matrices"""
        lb = None if labels is None else i32(labels)
        n = self.n if lb is None else len(lb)
        cnt = C.c_int64(0)
        rc = self.lib.stf_download_trees(self.h, n, ptr(lb), None, 0, C.byref(cnt))
        if rc not in (0, _lib.STF_E_CAPACITY):
            self._ck(rc)
        buf = np.zeros(max(1, cnt.value), np.uint8)
        self._ck(self.lib.stf_download_trees(self.h, n, ptr(lb), ptr(buf), len(buf), C.byref(cnt)))
        out, pos = [], 0
        for i in range(n):
            o = (int(lb[i]) if lb is not None else i + 1) - 1
            a, b, lv = int(self.kh1[o]), int(self.kw1[o]), []
            for _ in range(self.nlevels):
                lv.append(buf[pos:pos + a * b].reshape((a, b), order="F"))
                pos += a * b
                a, b = a // 2 + 1, b // 2 + 1
            out.append(lv)
        return out

    def getshifts(self, labels=None, level=1):
        n = self.n if labels is None else len(labels)
        lb = None if labels is None else i32(labels)
        rs, cs = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self._ck(self.lib.stf_get_shifts(self.h, n, ptr(lb), level, ptr(rs), ptr(cs)))
        return rs, cs

    def _shift(self, labels, level, s1, s2, mode):
        n = self.n if labels is None else len(labels)
        lb = None if labels is None else i32(labels)
        s1, s2 = i32(s1), i32(s2)
        assert len(s1) == n and len(s2) == n
        self._ck(self.lib.stf_set_shifts(self.h, n, ptr(lb), level, ptr(s1), ptr(s2), mode))

    def setshift(self, labels, level, s1, s2):
        self._shift(labels, level, s1, s2, 0)

    def shift(self, labels, level, s1, s2):
        self._shift(labels, level, s1, s2, 1)

    def build(self, labels=None, layer=2):
        lb = None if labels is None else i32(labels)
        self._ck(self.lib.stf_build(self.h, 0 if lb is None else len(lb), ptr(lb), layer))

    # results
    def _call_list(self, fn, *args):
        while True:
            cnt = C.c_int64(0)
            rc = fn(self.h, *args, ptr(self._buf), self._cap, C.byref(cnt))
            if rc == _lib.STF_E_CAPACITY:
                self._cap = int(cnt.value) + 1024
                self._buf = np.zeros(self._cap, ITEM_DT)
                rc = self.lib.stf_fetch_result(self.h, ptr(self._buf), self._cap, C.byref(cnt))
            self._ck(rc)
            return self._buf[: cnt.value].copy()

    def counters(self):
        out = np.zeros(8, np.int64)
        self._ck(self.lib.stf_counters(self.h, ptr(out)))
        return dict(zip(_lib.COUNTER_FIELDS, (int(x) for x in out)))

    def keep_items(self, on=True):
        """materialise the candidate list (`itemlist`) in the following many-body calls so that fetch_items() can return it"""
        self._ck(self.lib.stf_set_keep_items(self.h, int(bool(on))))
        return self

    def fetch_items(self):
        cnt = C.c_int64(0)
        rc = self.lib.stf_fetch_items(self.h, None, 0, C.byref(cnt))
        if rc not in (0, _lib.STF_E_CAPACITY):
            self._ck(rc)
        buf = np.zeros(max(1, cnt.value), ITEM_DT)
        self._ck(self.lib.stf_fetch_items(self.h, ptr(buf), len(buf), C.byref(cnt)))
        return buf[: cnt.value]


def items_to_list(arr):
    """structured array → [((i1, i2), (l, a, b)), ...] (Vector{CoItem})"""
    return [((int(r["i1"]), int(r["i2"])), (int(r["l"]), int(r["a"]), int(r["b"]))) for r in arr]


def items_to_array(items):
    if isinstance(items, np.ndarray):
        return np.ascontiguousarray(items, dtype=ITEM_DT)
    a = np.zeros(len(items), ITEM_DT)
    for k, ((i1, i2), (l, x, y)) in enumerate(items):
        a[k] = (i1, i2, l, x, y)
    return a


def qtrees(pics, mask=None, size=None, background=None, maskbackground=None, device=0) -> DeviceQTrees:
    """qtrees(pics; mask, background, maskbackground, size)  Stuffing.jl:48-63; the first tree is the mask"""
    ts = []
    if mask is not None:
        mq = maskqtree(mask, background=maskbackground)
        ts.append(mq)
        if size is None:
            size = mq.canvas
    elif size is None:
        size = 2 ** math.ceil(math.log2(max(max(np.shape(p)) for p in pics)))
    for p in pics:
        ts.append(qtree(p, size, background=background))
    return DeviceQTrees(ts, device=device)


def _dev(qts) -> DeviceQTrees:
    if isinstance(qts, DeviceQTrees):
        return qts
    qts = list(qts)
    if all(isinstance(t, HostTree) for t in qts):
        return DeviceQTrees(qts)
    raise TypeError("expected a DeviceQTrees or a list of HostTrees")


def _pair(q1, q2):
    """two trees → (device vector, label1, label2)"""
    if isinstance(q1, DeviceTree) and isinstance(q2, DeviceTree) and q1.owner is q2.owner:
        return q1.owner, q1.label, q2.label
    if isinstance(q1, HostTree) and isinstance(q2, HostTree):
        assert len(q1) == len(q2)  # qtree_functions.jl:45
        return DeviceQTrees([q1, q2]), 1, 2
    raise TypeError("both trees must live in the same DeviceQTrees (or both be HostTrees)")


# ------------------------------------------------------------------------------------------------ pair collision
def collision(q1, q2):
    """collision(Q1, Q2) → (l, a, b); l < 0: no collision  qtree_functions.jl:43-51"""
    d, i1, i2 = _pair(q1, q2)
    out = np.zeros(3, np.int32)
    d._ck(d.lib.stf_collision(d.h, i1, i2, ptr(out)))
    return tuple(int(x) for x in out)


def _pairs_at(d, fn, pairs, at):
    arr = np.zeros(len(pairs), ITEM_DT)
    for k, ((i1, i2), a) in enumerate(zip(pairs, at)):
        arr[k] = (i1, i2, a[0], a[1], a[2])
    out = np.zeros(len(pairs), ITEM_DT)
    d._ck(fn(d.h, len(arr), ptr(arr), ptr(out)))
    return [(int(r["l"]), int(r["a"]), int(r["b"])) for r in out]


def collision_dfs(q1, q2, at=None):
    """collision_dfs(Q1, Q2, i=(length(Q1),1,1))  qtree_functions.jl:2-20"""
    d, i1, i2 = _pair(q1, q2)
    return _pairs_at(d, d.lib.stf_collision_dfs, [(i1, i2)], [at or (d.nlevels, 1, 1)])[0]


def collision_randbfs(q1, q2, at=None):
    """_collision_randbfs(Q1, Q2, [at]) in canonical child order  qtree_functions.jl:21-42"""
    d, i1, i2 = _pair(q1, q2)
    return _pairs_at(d, d.lib.stf_collision_bfs, [(i1, i2)], [at or (d.nlevels, 1, 1)])[0]


def collision_bfs_items(qts, items):
    """every result (hit or miss) of the BFS for explicit (i1,i2)=>at items, in item order"""
    d = _dev(qts)
    arr = items_to_array(items)
    out = np.zeros(len(arr), ITEM_DT)
    d._ck(d.lib.stf_collision_bfs(d.h, len(arr), ptr(arr), ptr(out)))
    return out


# ------------------------------------------------------------------------------------------------ many-body
def totalcollisions_native_items(qts, coitems, raw=False):
    """_totalcollisions_native(qtrees, coitems)  qtree_functions.jl:89-110"""
    d = _dev(qts)
    arr = items_to_array(coitems)
    r = d._call_list(d.lib.stf_collide_items, len(arr), ptr(arr))
    return r if raw else items_to_list(r)


def totalcollisions_native(qts, labels=None, raw=False, **_kw):
    """totalcollisions_native(qtrees, labels=1:length(qtrees))  qtree_functions.jl:126-131"""
    if len(qts) <= 1:
        return np.zeros(0, ITEM_DT) if raw else []
    d = _dev(qts)
    lb = None if labels is None else i32(list(labels))
    r = d._call_list(d.lib.stf_totalcollisions_native, 0 if lb is None else len(lb), ptr(lb))
    return r if raw else items_to_list(r)


def totalcollisions_spacial(qts, labels=None, unique=True, raw=False, **_kw):
    """totalcollisions_spacial(qtrees[, labels]; unique)  qtree_functions.jl:225-263"""
    if len(qts) <= 1:
        return np.zeros(0, ITEM_DT) if raw else []
    d = _dev(qts)
    lb = None if labels is None else i32(list(labels))
    r = d._call_list(d.lib.stf_totalcollisions_spacial, 0 if lb is None else len(lb), ptr(lb), int(unique))
    return r if raw else items_to_list(r)


def totalcollisions(qts, labels=None, **kw):
    """dispatcher  qtree_functions.jl:270-276"""
    last = qts if labels is None else labels
    if len(last) > SPACIAL_ENABLE_THRESHOLD:
        return totalcollisions_spacial(qts, labels, **kw)
    return totalcollisions_native(qts, labels, **kw)


def _cells(buf):
    d = {}
    for r in buf:
        d.setdefault((int(r["l"]), int(r["a"]), int(r["b"])), []).append(int(r["label"]))
    return d


class HashSpacialQTree:
    """hash_spacial_qtree(): the result of locate!(qts[, labels], ·) as {cell: [labels]}  qtrees.jl:411-422"""

    def __init__(self):
        self.qtree = {}

    def tree(self):
        return self.qtree

    def empty(self):
        self.qtree = {}
        return self


def hash_spacial_qtree(qts=None):
    return HashSpacialQTree()


class LinkedSpacialQTree:
    """linked_spacial_qtree(qts): the persistent located-cell table of a DeviceQTrees  qtrees.jl:449-552"""

    def __init__(self, qts: DeviceQTrees):
        self.owner = qts
        if qts.n:
            qts._ck(qts.lib.stf_linked_reset(qts.h))

    def clear(self, label):  # clear!(t, label)
        lb = i32([label])
        self.owner._ck(self.owner.lib.stf_linked_clear(self.owner.h, 1, ptr(lb)))

    def collect_tree(self):
        d = self.owner
        cap = 4 * d.n + 4
        buf = np.zeros(cap, CELL_DT)
        cnt = C.c_int64(0)
        d._ck(d.lib.stf_linked_collect(d.h, ptr(buf), cap, C.byref(cnt)))
        return _cells(buf[: cnt.value])


def linked_spacial_qtree(qts):
    return LinkedSpacialQTree(_dev(qts))


def locate(qts, labels_or_tree=None, tree=None):
    """locate!(qts[, labels][, spqtree])  qtree_functions.jl:190-201"""
    d = _dev(qts)
    labels = None
    if isinstance(labels_or_tree, (HashSpacialQTree, LinkedSpacialQTree)):
        tree = labels_or_tree
    else:
        labels = labels_or_tree
    if tree is None:
        tree = HashSpacialQTree()
    lb = None if labels is None else i32(list(labels))
    nl = 0 if lb is None else len(lb)
    if isinstance(tree, LinkedSpacialQTree):
        assert tree.owner is d
        d._ck(d.lib.stf_linked_locate(d.h, nl, ptr(lb)))
        return tree
    cap = 4 * d.n + 4 if lb is None else 4 * nl + 4
    buf = np.zeros(cap, CELL_DT)
    cnt = C.c_int64(0)
    d._ck(d.lib.stf_locate(d.h, nl, ptr(lb), ptr(buf), cap, C.byref(cnt)))
    for k, v in _cells(buf[: cnt.value]).items():
        tree.qtree.setdefault(k, []).extend(v)
    return tree


def partialcollisions(qts, tree=None, labels=None, unique=True, raw=False, **_kw):
    """partialcollisions(qtrees, linkedspqtree, labels; unique)  qtree_functions.jl:277-330"""
    d = _dev(qts)
    if d.n == 0:
        return np.zeros(0, ITEM_DT) if raw else []
    if tree is None:
        tree = LinkedSpacialQTree(d)
    assert tree.owner is d
    lb = None if labels is None else i32(list(labels))
    r = d._call_list(d.lib.stf_partialcollisions, 0 if lb is None else len(lb), ptr(lb), int(unique))
    return r if raw else items_to_list(r)


def dynamiccollisions(qts, tree, moved: list, unlocated: list, **kw):
    """dynamiccollisions(qtrees, linkedspqtree, moved; unlocated)  qtree_functions.jl:340-356.
    `moved` / `unlocated` are ordered sets held as lists and updated in place."""
    d = _dev(qts)
    if len(moved) / len(d) > 0.6:
        r = totalcollisions(d, **kw)
        for m in moved:
            if m not in unlocated:
                unlocated.append(m)
    else:
        ms = set(moved)
        stale = [i for i in unlocated if i not in ms]
        if stale:
            locate(d, stale, tree)
        unlocated.clear()
        r = partialcollisions(d, tree, moved, **kw)
    moved.clear()
    seen = set()
    for (i1, i2), _ in (r if not isinstance(r, np.ndarray) else items_to_list(r)):
        for x in (i1, i2):
            if x not in seen:
                seen.add(x)
                moved.append(x)
    return r


class DynamicColliders:
    """DynamicColliders(qtrees)  qtree_functions.jl:357-370"""

    def __init__(self, qts):
        self.qtrees = _dev(qts)
        self.spqtree = LinkedSpacialQTree(self.qtrees)
        self.moved = list(range(1, len(self.qtrees) + 1))
        self.unlocated = []

    def union(self, c):  # union!(dc, c): ordered-set union, O(|moved| + |c|)
        seen = set(self.moved)
        for x in c:
            if x not in seen:
                seen.add(x)
                self.moved.append(x)

    def dynamiccollisions(self, **kw):
        return dynamiccollisions(self.qtrees, self.spqtree, self.moved, self.unlocated, **kw)


def getpositions(qts, inds=None):
    """getpositions(qtrees, inds)  Stuffing.jl:78-88: (x, y) of every object relative to the mask"""
    d = _dev(qts)
    rs, cs = d.getshifts()
    idx = range(1, d.n) if inds is None else [i + 1 for i in inds]
    return [(int(cs[i] - cs[0] + 1), int(rs[i] - rs[0] + 1)) for i in idx]


def setpositions(qts, inds, x_y):
    """setpositions!(qtrees, inds, x_y)  Stuffing.jl:89-99"""
    d = _dev(qts)
    rs, cs = d.getshifts([1])
    inds = list(range(d.n - 1)) if inds is None else list(inds)
    if isinstance(x_y, tuple):
        x_y = [x_y] * len(inds)
    labels = [i + 2 for i in inds]
    d.setshift(labels, 1, [p[1] - 1 + int(rs[0]) for p in x_y], [p[0] - 1 + int(cs[0]) for p in x_y])
    return x_y
