"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed over NCCL/NVLink).

The path shards two ways:
  * by SCENE — independent scenes never interact: ranks own disjoint scene ranges, no data-path collective;
  * by ITEM inside one scene (`ShardedFit`) — every rank holds an identical replica, generates and tests only the
    candidate pairs of the sorted records rank, rank+world, ...; per fit iteration ONE all-gather of the local hit
    lists (16 B per hit, fixed stride) is exchanged, then every rank runs the same canonical batchsteps, which keeps
    the replicas bit-identical without shipping pyramids.
torch.distributed is plumbing: the exchange buffers are torch tensors, everything else happens behind the C ABI
(stf_shard_collide / stf_shard_pack / stf_shard_step)."""
from __future__ import annotations

import ctypes as C

import numpy as np

HIT_WORDS = 4  # one exchanged record = 16 bytes = 4 x int32: (i1, i2 | l << 26, a, b); record 0 of a segment = header


def shard_range(n_units: int, rank: int, world: int) -> range:
    """contiguous, balanced range of scene (or item) indices owned by `rank`: the first n % world ranks get one more"""
    per, rem = divmod(n_units, world)
    lo = rank * per + min(rank, rem)
    return range(lo, lo + per + (1 if rank < rem else 0))


def reduce_job(ms_local: float, units_local: float, dist=None, device=None):
    """(max over ranks of the step time, sum over ranks of the units processed) → whole-job throughput inputs"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return ms_local, units_local
    import torch
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    u = torch.tensor([units_local], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


# ---- the exchange format, restated on the host (tests drive it over gloo; the device twins are k_shard_pack/unpack)
def exchange_stride(nlocal_max: int, quantum: int = 256) -> int:
    """records per rank segment: header + the longest local list, rounded up so that buffers are reused across rounds"""
    return ((nlocal_max + 1 + quantum - 1) // quantum) * quantum


def pack_segment(hits: np.ndarray, stride: int) -> np.ndarray:
    """hits: (n, 4) int32 records → one (stride, 4) segment: [count | records | padding]"""
    n = len(hits)
    if n + 1 > stride:
        raise ValueError("exchange stride too small for the local hit list")
    seg = np.zeros((stride, HIT_WORDS), np.int32)
    seg[0, 0] = n & 0x7FFFFFFF
    seg[0, 1] = n >> 31
    seg[1:n + 1] = hits
    return seg


def unpack_segments(recv: np.ndarray, world: int, stride: int) -> np.ndarray:
    """(world*stride, 4) gathered buffer → the union of the local lists in rank order"""
    recv = recv.reshape(world, stride, HIT_WORDS)
    parts = []
    for r in range(world):
        n = int(recv[r, 0, 0]) | (int(recv[r, 0, 1]) << 31)
        parts.append(recv[r, 1:n + 1])
    return np.concatenate(parts) if parts else np.zeros((0, HIT_WORDS), np.int32)


class ShardedFit:
    """item-parallel fit iterations of ONE scene replicated on every rank.  `d` is this rank's DeviceQTrees."""

    def __init__(self, d, rank: int = 0, world: int = 1, dist=None):
        import torch
        self.torch, self.d, self.rank, self.world, self.dist = torch, d, rank, world, dist
        self.dev = torch.device("cuda", d.device)
        self.send = self.recv = None
        self.bytes_exchanged = 0

    def _buffers(self, stride):
        t = self.torch
        if self.send is None or self.send.shape[0] < stride:
            self.send = t.zeros((stride, HIT_WORDS), dtype=t.int32, device=self.dev)
            self.recv = t.zeros((self.world * stride, HIT_WORDS), dtype=t.int32, device=self.dev) if self.world > 1 else self.send
        return self.send, self.recv

    def round(self, seed: int, iteration: int, labels=None, mode: int = 0) -> int:
        """one fit iteration; returns length(colist).  Collectives: one scalar all-reduce (stride) + ONE all-gather."""
        d, t = self.d, self.torch
        lb = None if labels is None else np.ascontiguousarray(labels, np.int32)
        nloc = C.c_int64(0)
        d._ck(d.lib.stf_shard_collide(d.h, mode, 0 if lb is None else len(lb), None if lb is None else lb.ctypes.data, self.rank, self.world, C.byref(nloc)))
        nmax = nloc.value
        if self.world > 1:
            m = t.tensor([nmax], dtype=t.int64, device=self.dev)
            self.dist.all_reduce(m, op=self.dist.ReduceOp.MAX)
            nmax = int(m.item())
        stride = exchange_stride(nmax)
        if self.send is not None and self.send.shape[0] >= stride:
            stride = self.send.shape[0]
        send, recv = self._buffers(stride)
        d._ck(d.lib.stf_shard_pack(d.h, send.data_ptr(), stride))
        if self.world > 1:
            self.dist.all_gather_into_tensor(recv, send)
            t.cuda.current_stream(self.dev).synchronize()
            self.bytes_exchanged += self.world * stride * 16
        nh = C.c_int64(0)
        d._ck(d.lib.stf_shard_step(d.h, recv.data_ptr(), self.world, stride, seed, iteration, C.byref(nh)))
        return nh.value


class GpuGroup:
    """Several GPUs driven from ONE process (stf_group_*): one context per device holding the same scene; fit rounds are
    split by candidate records and the GPUs exchange their hit lists themselves over NVLink peer memory (device-side flag
    barrier), so a round costs one host synchronisation per GPU and no collective call.  `replicas[i]` is the DeviceQTrees
    view of GPU i (state changes such as setshift must be applied to every replica: `each(fn)`)."""

    def __init__(self, trees, devices, is_mask=None, scene=None):
        from . import _lib
        from .qtrees import DeviceQTrees
        self.lib = _lib.load()
        devs = np.ascontiguousarray(np.asarray(list(devices), np.int32))
        self.h = C.c_void_p()
        rc = self.lib.stf_group_create(len(devs), devs.ctypes.data, C.byref(self.h))
        if rc != 0:
            raise _lib.StuffingError(rc, "stf_group_create failed (no CUDA device, or no peer access between the listed devices)")
        trees = list(trees)
        self.replicas = [DeviceQTrees(trees, device=int(dv), is_mask=is_mask, scene=scene, handle=self.lib.stf_group_ctx(self.h, i)) for i, dv in enumerate(devs)]
        self.devices = [int(x) for x in devs]

    def __len__(self):
        return len(self.replicas)

    def each(self, fn):
        return [fn(d) for d in self.replicas]

    def _ck(self, rc):
        if rc != 0:
            from . import _lib
            raise _lib.StuffingError(rc, self.lib.stf_group_last_error(self.h).decode())

    def setshift(self, labels, level, s1, s2, relative: bool = False):
        """setshift!/shift! of the listed trees on every replica (the GPUs driven in parallel)"""
        lb = np.ascontiguousarray(np.asarray(labels, np.int32)); s1 = np.ascontiguousarray(np.asarray(s1, np.int32)); s2 = np.ascontiguousarray(np.asarray(s2, np.int32))
        self._ck(self.lib.stf_group_set_shifts(self.h, len(lb), lb.ctypes.data, level, s1.ctypes.data, s2.ctypes.data, 1 if relative else 0))

    def reset_velocity(self):
        self._ck(self.lib.stf_group_reset_velocity(self.h, 0, None))

    def fit_round(self, seed: int, iteration: int, labels=None, partial: bool = False) -> int:
        lb = None if labels is None else np.ascontiguousarray(np.asarray(list(labels), np.int32))
        nh = C.c_int64(0)
        rc = self.lib.stf_group_fit_round(self.h, 1 if partial else 0, 0 if lb is None else len(lb), None if lb is None else lb.ctypes.data, seed, iteration, C.byref(nh))
        if rc != 0:
            from . import _lib
            raise _lib.StuffingError(rc, self.lib.stf_group_last_error(self.h).decode())
        return nh.value

    @property
    def nvlink_bytes(self) -> int:
        return int(self.lib.stf_group_nvlink_bytes(self.h))

    def close(self):
        if self.h:
            for d in self.replicas:
                d.h = None
            self.lib.stf_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
