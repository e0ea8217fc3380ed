"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/*.h declares, the
Python binding covers them, the product path refuses to run without a CUDA device, host-side helpers."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import ROOT, product


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "stuffing_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    S = product()
    lib = ctypes.CDLL(S._lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert getattr(lib, n) is not None, n
        assert n in S._lib.EXPORTS, f"{n} is declared in the header but missing from the Python binding"
    assert set(S._lib.EXPORTS) == set(names)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    S = product()
    with pytest.raises(S._lib.StuffingError):
        S.DeviceQTrees([S.qtree(np.ones((3, 4), bool))])
    # nothing under the product package imports the oracle
    for f in os.listdir(os.path.join(ROOT, "stuffing.jl_b200")):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "stuffing.jl_b200", f)).read().replace("the oracle", ""), f


def test_host_constructors_match_reference_conventions():
    S = product()
    t = S.qtree(np.ones((3, 4), bool), 16)            # test/runtests.jl:9-10
    assert t.codes[0, 0] == S.FULL and t.default == S.EMPTY and len(t) == 5
    t = S.qtree(np.ones((3, 4)), background=1.0)      # :12-13
    assert t.codes[0, 0] == S.EMPTY
    assert len(S.qtree(np.array([[1]]), background=0)) == 2          # test/test_qtrees.jl:22-25
    assert len(S.qtree(np.array([[1, 0]]), background=0)) == 2
    m = S.maskqtree(np.ones((500, 800), bool))        # Stuffing.jl:29-46: canvas 2^ceil(log2(ms+max(0.024ms,20))), centred
    assert m.canvas == 1024 and m.default == S.FULL and m.shift1 == ((1024 - 500) // 2, (1024 - 800) // 2)
    assert m.codes[0, 0] == S.EMPTY
    with pytest.raises(AssertionError):
        S.HostTree(np.ones((3, 3), np.uint8), 12)     # isinteger(log2(sz))  qtrees.jl:197
    x = np.random.default_rng(3).random(1000) * 1000
    assert sum(int(np.floor(np.log2(v))) for v in x) == sum(S.trainer.intlog2(v) for v in x)  # test/test_fit.jl:2-3
    lru = S.trainer.LRU()                             # test/test_fit.jl:5-13
    for i in range(1, 11):
        lru.push(i)
    for i in range(10, 1, -2):
        lru.push(i)
    lru.push(1)
    assert lru.collect() == [1, 2, 4, 6, 8, 10, 9, 7, 5, 3] and lru.collect(3) == [1, 2, 4]


def test_shard_range_partitions_units():
    S = product()
    from stuffing_b200 import parallel
    for n in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            parts = [parallel.shard_range(n, r, world) for r in range(world)]
            assert sum(len(p) for p in parts) == n
            assert [i for p in parts for i in p] == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["STF_ROOT"])
import torch, torch.distributed as dist
import _pkg
S = _pkg.load()
from stuffing_b200 import parallel
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
scenes = parallel.shard_range(4096, rank, world)
ms, units = parallel.reduce_job(10.0 + rank, float(len(scenes)), dist)
assert ms == 10.0 + world - 1 and units == 4096.0, (ms, units)
dist.barrier()
if rank == 0:
    print("OK", ms, units)
dist.destroy_process_group()
'''


def test_two_rank_gloo_reduction(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, STF_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29541", str(script)], env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK 11.0 4096.0" in out.stdout


_SHARD_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["STF_ROOT"])
import numpy as np, torch, torch.distributed as dist
import _pkg
S = _pkg.load()
from stuffing_b200 import parallel
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# every rank "finds" the hits of its share of the records: a deterministic global list, dealt round-robin
rng = np.random.default_rng(7)
allhits = rng.integers(1, 1000, size=(1237, 4)).astype(np.int32)
local = allhits[rank::world]
# the protocol of ShardedFit.round: scalar all-reduce(max) -> stride, fixed-stride all-gather, union in rank order
m = torch.tensor([len(local)], dtype=torch.int64); dist.all_reduce(m, op=dist.ReduceOp.MAX)
stride = parallel.exchange_stride(int(m.item()))
send = torch.from_numpy(parallel.pack_segment(local, stride))
recv = torch.zeros((world * stride, 4), dtype=torch.int32)
dist.all_gather_into_tensor(recv, send)
union = parallel.unpack_segments(recv.numpy(), world, stride)
want = np.concatenate([allhits[r::world] for r in range(world)])
assert np.array_equal(union, want), "union differs"
# every rank must hold the same union (the canonical sort + batchsteps then keep the replicas identical)
digest = torch.tensor([int(union.astype(np.int64).sum()), len(union)], dtype=torch.int64)
lo, hi = digest.clone(), digest.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert torch.equal(lo, hi)
try:
    parallel.pack_segment(local, len(local))      # stride must leave room for the header
    raise SystemExit("pack_segment accepted a short stride")
except ValueError:
    pass
if rank == 0:
    print("SHARD-OK", len(union), stride)
dist.destroy_process_group()
'''


def test_two_rank_gloo_shard_exchange(tmp_path):
    """host side of the item-parallel path: stride agreement, fixed-stride all-gather, union order"""
    script = tmp_path / "shard_worker.py"
    script.write_text(_SHARD_WORKER)
    env = dict(os.environ, STF_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29543", str(script)], env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "SHARD-OK 1237 768" in out.stdout
