"""torchrun check of the item-parallel path over NCCL (run under `gpurun --gpus N`):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 tools/check_shard.py
Every rank builds the same seeded scene twice: one context runs stf_fit_round alone (the single-GPU reference), the
other runs ShardedFit.round with the real all-gather.  After every iteration hit counts and all level-1 shifts must be
identical on every rank, bit for bit; prints per-iteration device times of both."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import _pkg, workloads
S = _pkg.load()
from stuffing_b200 import parallel

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
which = sys.argv[1] if len(sys.argv) > 1 else "c1"
fx = np.load(os.path.join(ROOT, "bench_data", "c4_rect_n100000_seed4.npz" if which == "c4" else "c1_word_n10000_seed1.npz"))
n, seed, side, canvas = int(fx["n"]), int(fx["seed"]), int(fx["mask_side"]), int(fx["canvas"])
rng = np.random.default_rng(seed)
objs = workloads.rect_objects(n, 30, rng) if which == "c4" else workloads.word_objects(n, rng)
trees = [S.maskqtree(np.ones((side, side), bool))] + [S.qtree(o, canvas) for o in objs]
for t, r, c in zip(trees, fx["rs"], fx["cs"]):
    t.shift1 = (int(r), int(c))
ref, rep = S.DeviceQTrees(trees, device=local), S.DeviceQTrees(trees, device=local)
for d in (ref, rep):
    d.lib.stf_set_optimiser(d.h, 2, 0.25, 0.5)
sh = parallel.ShardedFit(rep, rank, world, dist if world > 1 else None)
ok = True
for it in range(6):
    nh = C.c_int64(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ref._ck(ref.lib.stf_fit_round(ref.h, 0, 0, None, 9, it, C.byref(nh)))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    got = sh.round(9, it)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    r0, c0 = ref.getshifts(); r1, c1 = rep.getshifts()
    same = got == nh.value and np.array_equal(r0, r1) and np.array_equal(c0, c1)
    ok &= same
    if rank == 0:
        print(f"iter {it}: hits {nh.value} sharded {got} identical={same}  single {1e3*(t1-t0):.3f} ms  sharded x{world} {1e3*(t2-t1):.3f} ms", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
if rank == 0:
    print("SHARD-NCCL-OK" if int(flag.item()) else "SHARD-NCCL-MISMATCH", "world", world, which, flush=True)
sys.exit(0 if int(flag.item()) else 1)
