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


def _julia_ccalls():
    """(name, return type, [argument types]) of every `ccall((:name, LIB), Ret, (Args...), ...)` in julia/StuffingB200.jl"""
    import re
    src = open(os.path.join(ROOT, "julia", "StuffingB200.jl")).read()
    out = []
    for m in re.finditer(r"ccall\(\(:(\w+), LIB\),\s*([\w{}]+),\s*\(", src):
        i, depth = m.end(), 1
        while depth:                      # the argument-type tuple, with nested braces/parentheses
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        tup = src[m.end():i - 1]
        args, cur, lvl = [], "", 0
        for ch in tup:
            if ch == "," and lvl == 0:
                args.append(cur.strip()); cur = ""
            else:
                lvl += {"{": 1, "}": -1}.get(ch, 0); cur += ch
        if cur.strip():
            args.append(cur.strip())
        out.append((m.group(1), m.group(2), args))
    return out


def test_julia_glue_ccalls_match_the_library_exports():
    """the Julia binding cannot run here (no Julia in the image): every ccall in it is checked statically against the
    ctypes declarations of the same export — name, arity, and the C type class of the result and of every argument"""
    import ctypes as C
    S = product()
    jl2c = {"Int32": C.c_int32, "Int64": C.c_int64, "UInt64": C.c_uint64, "Float64": C.c_double, "Cstring": C.c_char_p, "Cvoid": None}
    calls = _julia_ccalls()
    assert len(calls) >= 30
    seen = set()
    for name, ret, args in calls:
        assert name in S._lib.EXPORTS, f"{name}: not an export of the library"
        cres, cargs = S._lib.EXPORTS[name]
        assert len(args) == len(cargs), (name, args, cargs)
        want_ret = C.c_void_p if ret.startswith("Ptr{") else jl2c[ret]
        assert want_ret == cres, (name, ret, cres)
        for a, ca in zip(args, cargs):
            if a.startswith("Ptr{"):
                assert ca in (C.c_void_p, C.c_char_p) or hasattr(ca, "contents") or "LP_" in getattr(ca, "__name__", ""), (name, a, ca)
            else:
                assert jl2c[a] == ca, (name, a, ca)
        seen.add(name)
    # the hot-path surface of SURVEY §8b is bound
    for must in ("stf_upload_objects", "stf_collision", "stf_totalcollisions_spacial", "stf_totalcollisions_native", "stf_partialcollisions", "stf_locate",
                 "stf_linked_locate", "stf_grad2d", "stf_batchsteps", "stf_fit_round", "stf_fit_rounds", "stf_set_shifts", "stf_get_shifts", "stf_download_trees",
                 "stf_group_create", "stf_group_fit_round", "stf_fit_round_positions", "stf_fit_round_moved", "stf_fit_epoch_e", "stf_filter_pairs", "stf_collide_items",
                 "stf_ground_compose", "stf_ground_level", "stf_shard_collide", "stf_shard_pack", "stf_shard_step", "stf_linked_clear", "stf_linked_collect",
                 "stf_get_velocity", "stf_set_level_shift", "stf_counters", "stf_group_set_shifts"):
        assert must in seen, must


def test_narrow_bit_helpers_on_the_host(tmp_path):
    """the pure bit helpers of the word-parallel narrow phase (stf_narrow.cuh: raster -> Morton order, frontier spreading,
    field squeezing) compiled for the HOST with nvcc and checked exhaustively / on random masks"""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    src = tmp_path / "t.cu"
    src.write_text(r'''
#include "%s/stuffing.jl_b200/csrc/stf_narrow.cuh"
#include <cstdio>
#include <cstdlib>
int main() {
    int bad = 0;
    for (int d = 1; d <= 3; d++) {
        int n = 1 << d;
        for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
            int be = n - 1 - j, al = n - 1 - i, key = 0;
            for (int t = d - 1; t >= 0; t--) key = (key << 2) | ((((be >> t) & 1) << 1) | ((al >> t) & 1));
            uint64_t r = 1ull << (j * n + i), m;
            if (d == 1) m = nw_morton1((uint32_t)r); else if (d == 2) m = nw_morton2((uint32_t)r); else m = nw_morton3(r);
            if (m != (1ull << key)) bad++;
            int a2, b2; nw_unkey((uint32_t)key, &a2, &b2);
            if (a2 != al || b2 != be) bad++;
        }
    }
    srand(1);
    for (int it = 0; it < 1000; it++) {
        uint64_t r = ((uint64_t)rand() << 40) ^ ((uint64_t)rand() << 20) ^ rand(), w = 0;
        for (int p = 0; p < 64; p++) if ((r >> p) & 1) w |= nw_morton3(1ull << p);
        if (w != nw_morton3(r)) bad++;
        uint32_t r16 = r & 0xFFFF, w16 = 0;
        for (int p = 0; p < 16; p++) if ((r16 >> p) & 1) w16 |= nw_morton2(1u << p);
        if (w16 != nw_morton2(r16)) bad++;
    }
    for (uint32_t m = 0; m < 16; m++) { uint32_t w = 0; for (int k = 0; k < 4; k++) if ((m >> k) & 1) w |= 0xFu << (4 * k); if (w != nw_children16(m)) bad++; }
    for (uint32_t m = 0; m < 65536; m++) { uint64_t w = 0; for (int k = 0; k < 16; k++) if ((m >> k) & 1) w |= 0xFull << (4 * k); if (w != nw_children64(m)) bad++; }
    for (int it = 0; it < 100000; it++) { uint32_t x = ((uint32_t)rand() << 16) ^ rand(), w = 0; for (int i = 0; i < 16; i++) if ((x >> (2 * i + 1)) & 1) w |= 1u << i; if (w != nw_squeeze_odd(x)) bad++; }
    printf("bad = %%d\n", bad);
    return bad != 0;
}
''' % ROOT)
    exe = tmp_path / "t"
    subprocess.check_call([nvcc, "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", str(exe), str(src)])
    assert subprocess.run([str(exe)], capture_output=True, text=True).returncode == 0


def test_charimage_and_npy_scene_exchange(tmp_path):
    """f-4 wire formats on the host: charimage glyphs and elision as src/qtrees.jl:308-345; save_scene/load_scene round
    trip through plain .npy files in Fortran order (what julia/NpyIO.jl reads)"""
    S = product()
    m = np.array([[1, 2, 3], [2, 2, 9]], np.uint8)
    assert S.charimage(m) == "░▓▒\n▓▓▞\n"
    big = np.full((60, 10), 1, np.uint8); big[0, 0] = 2
    rows = S.charimage(big, maxlen=9).split("\n")[:-1]
    assert len(rows) == 4 + 1 + 4 and rows[0][0] == "▓" and rows[4][0] == "⋮" and len(rows[4]) == 9   # 4 rows, an elision row, 4 rows
    wide = np.full((3, 60), 3, np.uint8)
    rows = S.charimage(wide, maxlen=9).split("\n")[:-1]
    assert all(len(r) == 9 and r[4] == "…" for r in rows)
    rng = np.random.default_rng(0)
    trees = [S.maskqtree(np.ones((40, 50), bool))] + [S.qtree(rng.random((int(rng.integers(3, 12)), int(rng.integers(3, 12)))) < 0.6, 128) for _ in range(5)]
    for k, t in enumerate(trees[1:]):
        t.shift1 = (3 * k + 1, 40 - 5 * k)
    S.save_scene(str(tmp_path / "scene"), trees)
    hdr = open(tmp_path / "scene" / "tree_00002.npy", "rb").read(128)
    assert hdr[:6] == b"\x93NUMPY" and b"'fortran_order': True" in hdr and b"|u1" in hdr
    back = S.load_scene(str(tmp_path / "scene"))
    assert len(back) == len(trees)
    for a, b in zip(trees, back):
        assert np.array_equal(a.codes, b.codes) and a.shift1 == b.shift1 and a.default == b.default and a.canvas == b.canvas
