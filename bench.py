#!/usr/bin/env python
"""bench.py — collision queries/s and fit iterations/s of the collision-and-fit hot path (BASELINE.json).

Workloads (`--workload`, one per BASELINE.json config; the default c1 is the configuration the metric is quoted on):
  c1  ONE scene of 10 000 word masks + the mask on a 4096^2 canvas.  One STEP = one fit iteration from a fixed seeded layout:
        restore: setshift!(t, 1, layout) of the objects 2..n the previous step moved (never the mask; a tree that already
                 sits there is left alone — both arms apply the same exact rule)            (subsystem 1: shift + rebuild)
        -> reset!(optimiser) -> totalcollisions_spacial(qtrees; unique=false)                (subsystems 2, 3)
        -> batchsteps!(qtrees, colist, Momentum(1/4, 1/2))                                   (subsystem 4)
      every step does identical work (same layout, same canonical draws) and both arms assert the fixture's item/hit counts.
  c2  examples/packing.jl-style fit loop: 1 000 rect masks sz 36 + a 1000^2 mask, canvas 1024; one STEP = restore +
      50 fixed rounds of (totalcollisions -> batchsteps!), draws keyed (seed 2, round).  Per-round hit counts and the final
      positions are asserted against the committed golden vector in both arms.
  c3  dynamiccollisions.jl-style loop on the c1 scene: per STEP 1 % of the trees jittered by +-(1..8) px, partialcollisions
      over (labels of the previous result + the jittered ones), batchsteps! on the result.  The state evolves.
  c4  100 000 rect masks + a 15900^2 mask, canvas 16384: the c1 step on the multi-launch throughput path.
  c5  batch of 4096 independent scenes x (mask + 128 word masks), canvas 512, one context: the c1 step for all scenes at once.

`value`  : narrow-phase items answered per second, whole job, inputs (label list + shift arrays) resident in HBM.
`e2e`    : the same through the C-ABI call with HOST buffers: shifts H2D every step, hit count and positions D2H every step.
`collide_only` : totalcollisions(qtrees[2:end]) alone, mask excluded, result list returned to the host — the call
           examples/collision_benchmark.jl:22-37 times (published: 21 244 us at n = 10 000 on an i7-8700, 4 threads).
`roofline` : the slowest stage of the step against the measured HBM peak (algorithmic bytes), every other
           stage under `stages`, and under `bandwidth_shaped` the kernels of the path that stream HBM (upload build), per kernel.
N > 1    : c1 / c2 / c3 / c4 — scene replicas ("weak"); c5 — the 4096 scenes are split over the ranks ("strong");
           `--shard items` — item-parallel sharding of ONE scene.
--impl reference : the CPU restatement of the reference algorithm (oracle/, kind "port": Julia is probed at run time and
           reported; the reference is pure Julia and has no C sources) on the host cores: same step, same metric, same
           config.workload string.
"""
import argparse
import ctypes as C
import gc
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
from workloads import load_workload, scene_batch  # noqa: E402

SEED, ITERATION = 1, 0
PUBLISHED_COLLIDE_MS = 21.24432552966102  # examples/collision_benchmark.jl:105 (n = 10 000; i7-8700, 4 threads, Julia 1.7.0)
L2_NOTE = ("GPU arm: L2 flushed between timed iterations (512 MiB write) for c1/c2/c4/c5; c3 not flushed (the state evolves from step to step, "
           "the 5 MB working set is L2-resident by design); CPU arm: not applicable")
NAMES = {"c1": "C1 fit iteration: 10k word masks + mask, 4096^2 canvas (restore moved trees, totalcollisions_spacial, batchsteps)",
         "c2": "C2 packing.jl fit loop: 1k rect masks + mask, 1024^2 canvas, 50 rounds of totalcollisions + batchsteps per step",
         "c3": "C3 dynamic steps (dynamiccollisions.jl style): 10k word masks + mask, 4096^2 canvas, partialcollisions over the moved subset + batchsteps",
         "c4": "C4 fit iteration: 100k rect masks + mask, 16384^2 canvas (restore moved trees, totalcollisions_spacial, batchsteps)",
         "c5": "C5 fit iteration: batch of 4096 scenes x (mask + 128 word masks), 512^2 canvas (restore moved trees, totalcollisions_spacial, batchsteps)"}


def julia_probe():
    """is the real reference runnable on this box?"""
    exe = shutil.which("julia")
    if not exe:
        return "absent"
    try:
        return subprocess.run([exe, "--version"], capture_output=True, text=True, timeout=20).stdout.strip() or "present"
    except Exception as e:
        return f"present but not runnable: {e}"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram__bytes_read+write per launch of the dominant kernels, from the committed ncu --set full captures"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.rows, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", os.environ.get("STF_BENCH_CLOCK_MS", "200")],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def c3_script(n_all, nsteps, seed=3):
    """the scripted moving subset of C3 (identical in both arms): per step 1 % of the trees and their +-(1..8) px jitter"""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(nsteps):
        jit = rng.choice(np.arange(2, n_all + 1), size=n_all // 100, replace=False).astype(np.int32)
        dr = (rng.integers(1, 9, len(jit)) * rng.choice([-1, 1], len(jit))).astype(np.int32)
        dc = (rng.integers(1, 9, len(jit)) * rng.choice([-1, 1], len(jit))).astype(np.int32)
        out.append((jit, dr, dc))
    return out


def labels_of_result(colist):
    """moved := the labels of the result except the mask (qtree_functions.jl:354-355), ascending (canonical order)"""
    if len(colist) == 0:
        return []
    u = np.union1d(colist["i1"], colist["i2"])
    return u[u != 1].astype(np.int32).tolist()


# ---------------------------------------------------------------------------------------------- reference arm (CPU)
class RefScene:
    """one single-scene workload on the CPU restatement; every method mirrors what the GPU arm does in the same step"""

    def __init__(self, orc, which, cores_all):
        self.orc, self.which = orc, which
        self.W = W = load_workload("c1" if which == "c3" else which)
        self.qts = workloads.oracle_scene(orc, W)
        self.n_all = len(self.qts)
        self.movable = np.arange(2, self.n_all + 1, dtype=np.int32)
        self.rs0, self.cs0 = np.ascontiguousarray(W["rs"][1:]), np.ascontiguousarray(W["cs"][1:])
        self.cores_all = cores_all
        self.cores = self.best_threads()
        self.rebuilt = 0

    def restore(self):
        self.rebuilt = self.orc.setshifts(self.qts, self.movable, 1, self.rs0, self.cs0, skip_same=True, nthreads=self.cores_all)

    def collide(self, labels=None, unique=False, nthreads=None):
        hits, info = self.orc.totalcollisions_spacial(self.qts, labels, unique=unique, nthreads=nthreads or self.cores, raw=True)
        if not unique:
            self.orc.lib().orc_sort_unique(hits.ctypes.data, C.byref(C.c_int64(len(hits))), 0)
        return hits, info

    def best_threads(self):
        """the thread count (1 or all cores) that runs the narrow phase faster on this host (shared vCPUs can make OpenMP
        slower than one thread)"""
        best, best_t = 1, None
        for nt in sorted({1, self.cores_all}):
            self.collide(nthreads=nt)
            t0 = time.perf_counter()
            self.collide(nthreads=nt)
            dt = time.perf_counter() - t0
            if best_t is None or dt < best_t:
                best, best_t = nt, dt
        return best

    def step_c1(self):
        self.restore()
        opt = self.orc.Opt("momentum", 0.25, 0.5, self.n_all)
        hits, info = self.collide()
        self.orc.batchsteps(self.qts, hits, opt, seed=SEED, iteration=ITERATION)
        return info["n_items"] + info["n_direct"], len(hits)

    def step_c2(self):
        W = self.W
        self.restore()
        opt = self.orc.Opt("momentum", 0.25, 0.5, self.n_all)
        items, round_hits = 0, []
        for k in range(W["rounds"]):
            hits, info = self.collide()
            items += info["n_items"] + info["n_direct"]; round_hits.append(len(hits))
            self.orc.batchsteps(self.qts, hits, opt, seed=W["seed"], iteration=k)
        assert round_hits == W["round_hits"].tolist(), "c2: per-round hit counts differ from the golden vector"
        return items, sum(round_hits)

    def collide_only(self, reps=5):
        """totalcollisions(qts) with the mask excluded (examples/collision_benchmark.jl:22-37), result list included"""
        self.restore()
        ts = []
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            hits, info = self.collide(self.movable, unique=True)
            ts.append(time.perf_counter() - t0)
        ts = ts[1:]
        return {"ms_min": min(ts) * 1e3, "ms_mean": float(np.mean(ts)) * 1e3, "items": info["n_items"] + info["n_direct"], "collisions": len(hits), "threads": self.cores}


def run_reference(args):
    """times the CPU restatement on rank 0's host cores; c5 runs a bounded sample of the batch (scene by scene)"""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import oracle as orc  # the one place bench.py executes oracle/: the CPU baseline
    cores_all = os.cpu_count() or 1
    which = args.workload
    extra = {}
    if which == "c5":
        import _pkg
        S = _pkg.load()
        nsc = 8
        B = scene_batch(S, range(nsc))
        per = len(B["kh"]) // nsc
        scenes, pos = [], 0
        for s in range(nsc):
            ts = []
            for k in range(per):
                i = s * per + k
                n = int(B["kh"][i]) * int(B["kw"][i])
                ts.append(orc.OTree(B["payload"][pos:pos + n].reshape((int(B["kh"][i]), int(B["kw"][i])), order="F"), B["canvas"], int(B["dflt"][i])))
                pos += n
            ts = orc.TreeList(ts)
            scenes.append((ts, np.ascontiguousarray(B["rs"][s * per + 1:(s + 1) * per]), np.ascontiguousarray(B["cs"][s * per + 1:(s + 1) * per])))
            orc.setshifts(ts, None, 1, B["rs"][s * per:(s + 1) * per], B["cs"][s * per:(s + 1) * per], skip_same=False)
        cores = 1
        mov = np.arange(2, per + 1, dtype=np.int32)

        def step():
            items = nh = 0
            for ts, rs, cs in scenes:
                orc.setshifts(ts, mov, 1, rs, cs, skip_same=True)
                opt = orc.Opt("momentum", 0.25, 0.5, per)
                hits, info = orc.totalcollisions_spacial(ts, unique=False, raw=True)
                orc.lib().orc_sort_unique(hits.ctypes.data, C.byref(C.c_int64(len(hits))), 0)
                orc.batchsteps(ts, hits, opt, seed=SEED, iteration=ITERATION)
                items += info["n_items"] + info["n_direct"]; nh += len(hits)
            return items, nh
        sample = f"{nsc} of the 4096 scenes per step, scene after scene on one core"
    elif which == "c3":
        R = RefScene(orc, "c3", cores_all)
        cores = R.cores
        tree = orc.LinkedTree(R.qts)
        tree.locate(R.qts)
        args.warmup = max(3, args.warmup)  # the state evolves: both arms time the same steps of the script
        script = c3_script(R.n_all, args.warmup + args.steps)
        state = dict(moved=R.movable.tolist(), k=0, items=0, hits=0)

        def step():
            jit, dr, dc = script[state["k"]]
            for lb, a, b in zip(jit.tolist(), dr.tolist(), dc.tolist()):
                R.qts[lb - 1].shift(1, a, b)
            labels = list(dict.fromkeys(state["moved"] + jit.tolist()))
            colist, info = orc.partialcollisions(R.qts, tree, labels, unique=False, nthreads=cores, raw=True)
            colist = colist.copy()
            orc.lib().orc_sort_unique(colist.ctypes.data, C.byref(C.c_int64(len(colist))), 0)
            opt_ = state.setdefault("opt", orc.Opt("momentum", 0.25, 0.5, R.n_all))
            orc.batchsteps(R.qts, colist, opt_, seed=SEED, iteration=state["k"])
            state["moved"] = labels_of_result(colist)
            state["k"] += 1
            return info["n_items"] + info["n_direct"], len(colist)
        sample = "the same scripted dynamic steps (jitter 1 % of the trees, partialcollisions over the moved subset, batchsteps)"
    else:
        R = RefScene(orc, which, cores_all)
        cores = R.cores
        step = R.step_c2 if which == "c2" else R.step_c1
        sample = "full steps of the C restatement (OpenMP over the narrow-phase items and over the restored trees)"
        if which in ("c1", "c4"):
            step()
            extra["collide_only"] = R.collide_only()
            extra["collide_only"]["published_ms"] = PUBLISHED_COLLIDE_MS if which == "c1" else None
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    items = hits = 0
    for _ in range(args.steps):
        a, b = step()
        items += a; hits += b
    ms = (time.perf_counter() - t0) / args.steps * 1e3
    items /= args.steps; hits /= args.steps
    if which in ("c1", "c4"):
        assert items == R.W["n_items"] and hits == R.W["n_hits"], (items, hits, "differ from the fixture's oracle counts")
        extra["trees_rebuilt_by_restore_per_step"] = int(R.rebuilt)
    v = items / (ms * 1e-3)
    rounds = 50 if which == "c2" else 1
    line = {"impl": "reference", "metric": "collision_queries_per_sec", "value": v, "unit": "items/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "fit_iters_per_sec": rounds * 1e3 / ms, "julia": julia_probe(),
            "config": {"workload": NAMES[which], "l2": L2_NOTE}, "items_per_step": int(items), "hits_per_step": int(hits),
            "cpu_baseline": {"value": v, "unit": "items/s", "cores": cores, "kind": "port", "sample": f"{args.steps} x {sample}"},
            "e2e": {"value": v, "unit": "items/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line.update(extra)
    print(json.dumps(line), flush=True)


def cpu_baseline(which, budget_s=20.0):
    """the oracle port timed on the host cores for a bounded sample of the same workload (single scenes: full steps)"""
    try:
        from oracle import oracle as orc
        if which == "c5":
            return {"value": None, "unit": "items/s", "cores": 0, "kind": "port", "sample": "run `bench.py --impl reference --workload c5` (8 of the 4096 scenes per step)"}
        R = RefScene(orc, which, os.cpu_count() or 1)
        fn = R.step_c2 if which == "c2" else R.step_c1
        if which == "c3":
            return {"value": None, "unit": "items/s", "cores": R.cores, "kind": "port", "sample": "run `bench.py --impl reference --workload c3` (the same scripted dynamic steps)"}
        fn()
        ts, items, t_all = [], 0, time.perf_counter()
        while len(ts) < 3 or (len(ts) < 10 and time.perf_counter() - t_all < budget_s):
            t0 = time.perf_counter()
            items, _h = fn()
            ts.append(time.perf_counter() - t0)
        best = min(ts)
        out = {"value": items / best, "unit": "items/s", "cores": R.cores, "kind": "port",
               "sample": f"{len(ts)} full steps of the same scene (restore moved trees + totalcollisions_spacial + sort + batchsteps), best", "ms_per_step": best * 1e3,
               "trees_rebuilt_by_restore_per_step": int(R.rebuilt)}
        if which in ("c1", "c4"):
            out["collide_only"] = R.collide_only(3)
        return out
    except Exception as e:  # the baseline is reporting, never the product path
        return {"value": None, "unit": "items/s", "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}


# ---------------------------------------------------------------------------------------------- our arm (GPU)
def build_roofline(S, torch, local, peak, scenes=4096):
    """the bandwidth-shaped kernels of the path: fused level-1 pack + pyramid build of a C5 batch whose u8 payload is
    resident in HBM (1.5 GB >> L2), reported PER KERNEL (stf_profile_read_kernels: an event pair around each launch):
    small = the 128 word masks of every scene, mid512 = the 490^2 masks.  Algorithmic bytes: the u8 read +
    0.25 B per level-1 node written + 5 x 0.25 B per upper node (4 children read, 1 written)."""
    B = scene_batch(S, range(scenes))
    dev = torch.from_numpy(np.concatenate([B["payload"], np.zeros(64, np.uint8)])).cuda(local)
    lib = S._lib.load()
    h = C.c_void_p()
    assert lib.stf_create(local, C.byref(h)) == 0
    lib.stf_profile_enable(h, 1)
    L = int(np.log2(B["canvas"])) + 1

    def alg_bytes(sel):
        a, b, nodes = B["kh"][sel].astype(np.int64), B["kw"][sel].astype(np.int64), 0
        payload = int((a * b).sum())
        for l in range(1, L + 1):
            nodes += int((a * b).sum()) * (1 if l == 1 else 5)
            a, b = a // 2 + 1, b // 2 + 1
        return float(payload) + 0.25 * nodes
    is_m = B["is_mask"] == 1
    alg_all, alg_small, alg_mask = alg_bytes(slice(None)), alg_bytes(~is_m), alg_bytes(is_m)
    ms = np.zeros(9); calls = np.zeros(9, np.int64)
    kms = np.zeros(4)
    best, best_k = None, None
    for _ in range(4):
        rc = lib.stf_upload_packed(h, len(B["kh"]), C.c_void_p(dev.data_ptr()), 1, B["kh"].ctypes.data, B["kw"].ctypes.data, B["rs"].ctypes.data, B["cs"].ctypes.data,
                                   B["dflt"].ctypes.data, B["canvas"], B["is_mask"].ctypes.data, B["scene"].ctypes.data)
        assert rc == 0, lib.stf_last_error(h)
        lib.stf_profile_read(h, ms.ctypes.data, calls.ctypes.data, 1)
        if hasattr(lib, "stf_profile_build_kernels"):
            lib.stf_profile_build_kernels(h, kms.ctypes.data)
        if best is None or ms[8] < best:
            best, best_k = float(ms[8]), kms.copy()
    lib.stf_destroy(h)
    ach = alg_all / (best * 1e-3) / 1e9
    out = {"kernel": "k_pack_build_small + k_pack_build_mid<512> (upload build of %d C5 scenes, %d trees)" % (scenes, len(B["kh"])), "bound": "hbm", "achieved": ach, "peak": peak,
           "unit": "GB/s", "frac": ach / peak, "ms": best, "algorithmic_bytes": int(alg_all), "payload_bytes": int(B["payload"].size), "input": "u8 payload resident in HBM, larger than L2"}
    if best_k is not None and best_k[0] > 0:
        per = {}
        for name, t, a in (("k_pack_build_small", best_k[0], alg_small), ("k_pack_build_mid<512>", best_k[2], alg_mask)):
            if t > 0:
                g = a / (t * 1e-3) / 1e9
                per[name] = {"ms": float(t), "algorithmic_bytes": int(a), "achieved": g, "frac": g / peak}
        out["per_kernel"] = per
        out["note"] = "per_kernel: each launch timed alone by its own event pair (they run back to back on one stream); the blended figure divides all bytes by the whole build stage"
    return out


class DevScene:
    """the GPU arm of one workload: upload + the step in its resident and its end-to-end form"""

    def __init__(self, S, torch, which, local, rank, world, shard_items, dist):
        from stuffing_b200 import parallel
        self.S, self.which, self.torch = S, which, torch
        t_up = time.perf_counter()
        if which == "c5":
            my = parallel.shard_range(int(os.environ.get("STF_BENCH_C5_SCENES", "4096")), rank, world)  # (env: a rank's share of a larger split, for analysis)
            B = scene_batch(S, my)
            self.d = S.DeviceQTrees.from_packed(B["payload"], B["kh"], B["kw"], B["rs"], B["cs"], B["dflt"], B["canvas"], B["is_mask"], B["scene"], device=local)
            rs0, cs0 = B["rs"], B["cs"]
            self.movable = np.flatnonzero(B["is_mask"] == 0).astype(np.int32) + 1
            self.shapes = list(zip(B["kh"][self.movable - 1].tolist(), B["kw"][self.movable - 1].tolist()))
            self.scaling, self.par, self.W = "strong", f"scene-parallel: {len(my)} of 4096 scenes per rank x{world}", None
        else:
            self.W = W = load_workload("c1" if which == "c3" else which)
            self.d = workloads.device_scene(S, W, local)
            rs0, cs0 = W["rs"], W["cs"]
            self.movable = np.arange(2, len(self.d) + 1, dtype=np.int32)  # the mask (label 1) never moves: restore objects 2..n only
            self.shapes = [o.shape for o in W["objs"]]
            self.scaling = "strong" if shard_items else "weak"
            self.par = f"item-parallel shards of one scene x{world} (all-gather of hit records per step)" if shard_items else f"scene replicas x{world}"
        self.t_up = time.perf_counter() - t_up
        d = self.d
        self.lib, self.h = d.lib, d.h
        self.n_all, self.nm = len(d), len(self.movable)
        self.rs0, self.cs0 = np.asarray(rs0), np.asarray(cs0)
        self.stream = torch.cuda.ExternalStream(self.lib.stf_stream(self.h), device=torch.device("cuda", local))
        self.rs_h = torch.from_numpy(np.ascontiguousarray(self.rs0[self.movable - 1])).pin_memory()
        self.cs_h = torch.from_numpy(np.ascontiguousarray(self.cs0[self.movable - 1])).pin_memory()
        self.rs_d, self.cs_d = self.rs_h.cuda(), self.cs_h.cuda()
        self.movable_d = torch.from_numpy(self.movable).cuda()  # resident runs: the label list lives in HBM like the positions
        self.out_rs = torch.empty(self.n_all, dtype=torch.int32).pin_memory(); self.out_cs = torch.empty(self.n_all, dtype=torch.int32).pin_memory()
        self.lib.stf_set_optimiser(self.h, 2, 0.25, 0.5)
        self.nh = C.c_int64(0)
        self.sharded = None
        self.rounds = int(self.W["rounds"]) if which == "c2" else 1
        self.round_hits = np.zeros(max(1, self.rounds), np.int64)
        self.tot = {}

    def restore(self, resident):
        d, lib, h = self.d, self.lib, self.h
        if resident:
            d._ck(lib.stf_set_shifts(h, self.nm, self.movable_d.data_ptr(), 1, self.rs_d.data_ptr(), self.cs_d.data_ptr(), 0 | 2 | 4))
        else:
            d._ck(lib.stf_set_shifts(h, self.nm, self.movable.ctypes.data, 1, self.rs_h.data_ptr(), self.cs_h.data_ptr(), 0))
        d._ck(lib.stf_reset_velocity(h, 0, None))

    def step(self, resident, count=False):
        d, lib, h, nh = self.d, self.lib, self.h, self.nh
        self.restore(resident)
        if self.which == "c2":  # 50 rounds, ONE call and one host synchronisation (stf_fit_rounds)
            tot8 = np.zeros(8, np.int64)
            d._ck(lib.stf_fit_rounds(h, 0, 0, None, int(self.W["seed"]), 0, self.rounds, self.round_hits.ctypes.data, tot8.ctypes.data,
                                     0 if resident else 1, None if resident else self.out_rs.data_ptr(), None if resident else self.out_cs.data_ptr()))
            if count:
                t8 = dict(zip(self.S._lib.COUNTER_FIELDS, (int(x) for x in tot8)))
                self.tot = dict(items=t8["items"], hits=t8["hits"], visited=t8["visited"], moved=t8["moved"], records=t8["records"] * self.rounds)
            return int(self.round_hits.sum())
        if self.sharded is not None:
            nh.value = self.sharded.round(SEED, ITERATION)
            if not resident:
                d._ck(lib.stf_get_shifts(h, self.n_all, None, 1, self.out_rs.data_ptr(), self.out_cs.data_ptr()))
        elif resident:
            d._ck(lib.stf_fit_round(h, 0, 0, None, SEED, ITERATION, C.byref(nh)))
        else:  # the user-facing call: the round and the positions after it
            d._ck(lib.stf_fit_round_positions(h, 0, 0, None, SEED, ITERATION, C.byref(nh), 1, self.out_rs.data_ptr(), self.out_cs.data_ptr()))
        if count:
            c = d.counters()
            self.tot = dict(items=c["items"] + c["direct"], hits=nh.value, visited=c["visited"], moved=c["moved"], records=c["records"])
        return nh.value

    def collide_only(self, reps=10):
        """totalcollisions(qtrees[2:end]) (mask excluded), the result list delivered to a host buffer"""
        torch, lib, h, d = self.torch, self.lib, self.h, self.d
        self.restore(True)
        buf = np.zeros(1 << 16, self.S._lib.ITEM_DT); cnt = C.c_int64(0)
        ts = []
        for _ in range(reps + 2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(self.stream)
            d._ck(lib.stf_totalcollisions_spacial(h, self.nm, self.movable.ctypes.data, 1, buf.ctypes.data, len(buf), C.byref(cnt)))
            b.record(self.stream); b.synchronize()
            ts.append(a.elapsed_time(b))
        ts = ts[2:]
        c = d.counters()
        out = {"ms_min": min(ts), "ms_mean": float(np.mean(ts)), "items": c["items"] + c["direct"], "collisions": int(cnt.value),
               "h2d_bytes": int(4 * self.nm), "d2h_bytes": int(20 * cnt.value + 136)}
        if self.which == "c1":
            out["published_ms"] = PUBLISHED_COLLIDE_MS
            out["vs_published"] = PUBLISHED_COLLIDE_MS / out["ms_mean"]
            out["note"] = "published: examples/collision_benchmark.jl:105 (i7-8700, 4 threads, Julia 1.7.0, WordCloud.rendertext words); here: synthetic word-like masks of the same size law, same n, same mask"
        return out


def run_ours(args):
    import torch
    import _pkg
    S = _pkg.load()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    which = args.workload
    shard_items = False  # (`--shard items` is run_group: one process drives all GPUs)
    sc = DevScene(S, torch, which, local, rank, world, shard_items, dist)
    d, lib, h = sc.d, sc.lib, sc.h
    stream = sc.stream
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def timed(resident, steps):
        tot = 0.0
        gc.collect(); gc.disable()  # (a collector pause on one rank is a straggler in the max over ranks)
        try:
            return _timed(resident, steps)
        finally:
            gc.enable()

    def _timed(resident, steps):
        tot = 0.0
        for _ in range(steps):
            if not os.environ.get("STF_BENCH_NOFLUSH"):  # (experiments only: the reported lines always flush)
                with torch.cuda.stream(stream):
                    flush.zero_()  # L2 flush between timed iterations, outside the event pair
            if sc.sharded is not None:
                torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            sc.step(resident)
            b.record(stream)
            b.synchronize()
            tot += a.elapsed_time(b)
        return tot

    def bracket():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(3, args.warmup)
    for _ in range(warm):
        sc.step(True); sc.step(False)
    sc.step(True, count=True)
    T = dict(sc.tot)
    items, hits = T["items"], T["hits"]
    # ---- the work is the fixture's: same counts as the oracle computed when the layout was generated
    if which in ("c1", "c4") and sc.sharded is None:
        assert items == sc.W["n_items"] and hits == sc.W["n_hits"], (items, hits, sc.W["n_items"], sc.W["n_hits"], "differ from the fixture's oracle counts")
    if which == "c2":
        sc.step(False)
        assert sc.round_hits.tolist() == sc.W["round_hits"].tolist(), "c2: per-round hit counts differ from the golden vector"
        assert np.array_equal(sc.out_rs.numpy(), sc.W["final_rs"]) and np.array_equal(sc.out_cs.numpy(), sc.W["final_cs"]), "c2: fitted positions differ from the golden vector"
        assert items == int(sc.W["round_items"].sum())

    sampler = ClockSampler(local) if rank == 0 and not os.environ.get("STF_BENCH_NOSAMPLER") else None
    l0 = lib.stf_launch_count(h)
    bracket()
    ms_res = timed(True, args.steps)
    bracket()
    launches = lib.stf_launch_count(h) - l0
    ms_e2e = timed(False, args.steps)
    bracket()
    # the per-stage split comes from a separate pass over the same steps: the stage timers (an event pair per stage)
    # stay out of the two timed regions above
    lib.stf_profile_enable(h, 1)
    lib.stf_profile_read(h, None, None, 1)
    timed(True, args.steps)
    stage_ms = np.zeros(9); stage_calls = np.zeros(9, np.int64)
    lib.stf_profile_read(h, stage_ms.ctypes.data, stage_calls.ctypes.data, 1)
    lib.stf_profile_enable(h, 0)
    bracket()
    clocks = sampler.stop() if sampler else None
    sc.step(False, count=True)
    if sc.sharded is None:
        assert sc.tot["items"] == items and sc.tot["hits"] == hits, "steps must do identical work"

    if os.environ.get("STF_BENCH_DEBUG_RANKS"):
        print(f"[rank {rank}] resident {ms_res / args.steps:.4f} ms  e2e {ms_e2e / args.steps:.4f} ms", file=sys.stderr, flush=True)
    t = torch.tensor([ms_res, ms_e2e, float(items), float(hits), float(T["visited"])], dtype=torch.float64, device="cuda")
    tmax = t.clone()
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms_res, ms_e2e = (float(x) / args.steps for x in tmax[:2].tolist())
    if which == "c5" or shard_items:      # every rank owns a disjoint part of the job: units add up
        items_job, hits_job = float(t[2]), float(t[3])
        if shard_items:
            hits_job = float(hits)        # the hit list is the union, identical on every rank
    else:                                 # replicas: N copies of the same job
        items_job, hits_job = float(items) * world, float(hits) * world
    # ---- the sharded configurations of the path, next to the replica headline: (1) C5, the 4096 scenes split
    # over the ranks (strong scaling, no data-path collective); (2) C1 item-parallel: rank 0 drives ALL GPUs of the box from
    # one process through stf_group_* (hit lists exchanged by the GPUs over NVLink peer memory).  Both are timed at every N,
    # N = 1 included, so the driver's per-N files carry the speed-ups.
    sharded = None
    if which == "c1" and not args.no_sharded:
        sharded = {}
        try:
            sc5 = DevScene(S, torch, "c5", local, rank, world, False, dist)
            for _ in range(3):
                sc5.step(True)
            sc5.step(True, count=True)
            bracket()
            tot5 = 0.0
            for _ in range(5):
                with torch.cuda.stream(sc5.stream):
                    flush.zero_()
                a5, b5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a5.record(sc5.stream); sc5.step(True); b5.record(sc5.stream); b5.synchronize()
                tot5 += a5.elapsed_time(b5)
            t5 = torch.tensor([tot5 / 5, float(sc5.tot["items"])], dtype=torch.float64, device="cuda")
            t5max = t5.clone()
            if dist is not None:
                dist.all_reduce(t5max, op=dist.ReduceOp.MAX); dist.all_reduce(t5, op=dist.ReduceOp.SUM)
            sharded["c5_scene_split"] = {"workload": NAMES["c5"], "gpus": world, "scaling": "strong", "ms_per_step": float(t5max[0]), "items_per_step": int(float(t5[1])),
                                         "items_per_s": float(t5[1]) / (float(t5max[0]) * 1e-3), "scene_iterations_per_s": 4096e3 / float(t5max[0]), "collective": "none (independent scenes)"}
            del sc5
        except Exception as e:  # reporting only
            sharded["c5_scene_split"] = {"error": str(e)}
        bracket()
        # rank 0 now drives every GPU of the box from its own process; the other ranks must leave their GPUs ALONE meanwhile
        # (an NCCL barrier would park a spinning kernel on each of them and time-slice against rank 0's work): they wait in a
        # host-side (gloo) barrier
        host_group = dist.new_group(backend="gloo") if dist is not None else None
        if rank == 0:
            try:
                ndev = min(world, torch.cuda.device_count())
                sharded["c1_item_parallel_group"] = group_bench(S, torch, "c1", list(range(ndev)), 10, 3)
            except Exception as e:
                sharded["c1_item_parallel_group"] = {"error": str(e)}
        if host_group is not None:
            dist.barrier(group=host_group)
        bracket()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline (SURVEY §8d algorithmic bytes): every stage, the slowest one on top
    per_step = stage_ms[:8] / args.steps
    names = S._lib.STAGES
    L = d.nlevels
    # the restore rebuilds only the trees the previous step moved (stf_set_shifts leaves a tree that already sits at the
    # requested shift alone): count those, from the positions the last end-to-end step read back
    mv = sc.movable - 1
    moved_sel = (sc.out_rs.numpy()[mv] != sc.rs0[mv]) | (sc.out_cs.numpy()[mv] != sc.cs0[mv])
    up_nodes = float(sum(5 * 0.25 * ((((kh - 2) >> (l - 1)) + 2) * (((kw - 2) >> (l - 1)) + 2))
                         for (kh, kw), m in zip(sc.shapes, moved_sel.tolist()) if m for l in range(2, L + 1)))
    R = sc.rounds
    alg = {  # algorithmic bytes per step for each stage (this rank)
        "shift_rebuild": up_nodes + 8.0 * sc.nm,
        "locate": R * (16.0 * sc.n_all + 0.25 * 4 * 6 * sc.n_all),
        "sort": 2 * 12.0 * T["records"],
        "emit": 16.0 * items,
        "narrow": 16.0 * items + 0.5 * T["visited"] + 16.0 * hits,
        "result": 2 * 16.0 * hits,
        "step_prep": 2 * 12.0 * 2 * hits,
        "step": hits * (2 * 8 * 0.25 + 2 * (8 + 16)) + 1.25 * 60 * T["moved"],
    }
    peak, peak_src = peaks()
    traffic = ncu_traffic()
    stages = {}
    for i in range(8):
        if per_step[i] <= 0:
            continue
        a = alg[names[i]] / (per_step[i] * 1e-3) / 1e9
        stages[names[i]] = {"ms": round(float(per_step[i]), 4), "algorithmic_bytes": int(alg[names[i]]), "achieved": a, "frac": a / peak}
    dom = names[int(np.argmax(per_step))]
    total_alg = sum(alg.values())
    roof = {"bound": "hbm", "kernel": dom, "achieved": stages[dom]["achieved"], "peak": peak, "unit": "GB/s", "frac": stages[dom]["frac"],
            "traffic": traffic.get(which, {}).get(dom), "peak_source": peak_src, "stages": stages,
            "whole_step": {"algorithmic_bytes": int(total_alg), "achieved": total_alg / (ms_res * 1e-3) / 1e9, "frac": total_alg / (ms_res * 1e-3) / 1e9 / peak},
            "note": "single-scene rounds are latency-bound (the whole step moves ~15 MB); the HBM-streaming kernels of the path are under bandwidth_shaped"}
    collide_only = None
    if which in ("c1", "c4") and sc.sharded is None:
        collide_only = sc.collide_only()
    if not args.no_build_roofline:
        try:
            roof["bandwidth_shaped"] = build_roofline(S, torch, local, peak)
            roof["bandwidth_shaped"]["traffic"] = traffic.get("build")
        except Exception as e:  # reporting only
            roof["bandwidth_shaped"] = {"error": str(e)}

    cpu = cpu_baseline(which) if world == 1 else {"value": None, "unit": "items/s", "cores": 0, "kind": "port", "sample": "timed at N=1 only"}
    scenes_job = 4096 if which == "c5" else (1 if shard_items else world)
    line = {"metric": "collision_queries_per_sec", "value": items_job / (ms_res * 1e-3), "unit": "items/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_res, "higher_is_better": True, "scaling": sc.scaling, "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "fit_iters_per_sec": scenes_job * R * 1e3 / ms_res, "julia": julia_probe(),
            "config": {"workload": NAMES[which], "l2": L2_NOTE},
            "details": {"objects": sc.n_all, "levels": L, "items_per_step": int(items_job), "hits_per_step": int(hits_job), "rounds_per_step": R,
                        "trees_rebuilt_by_restore_per_step": int(moved_sel.sum()),
                        "visited_nodes_per_step": int(float(t[4])) if which == "c5" or shard_items else T["visited"] * world,
                        "parallelism": sc.par, "upload_build_s": round(sc.t_up, 3),
                        "counts_asserted": "items/hits == fixture (oracle) counts" if which in ("c1", "c4") else ("per-round hits and final positions == golden vector" if which == "c2" else "steps do identical work")},
            "e2e": {"value": items_job / (ms_e2e * 1e-3), "unit": "items/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(12 * sc.nm), "d2h_bytes_per_step": int(8 * sc.n_all + 8 * R + 136 + (32 * (R + 2) if which == "c2" else 0))},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    if collide_only:
        line["collide_only"] = collide_only
    if sharded:
        line["sharded"] = sharded
    if sc.sharded is not None:
        line["nvlink_bytes_per_step"] = int(sc.sharded.bytes_exchanged / max(1, 2 * (args.steps + warm)))
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def group_bench(S, torch, which, devices, steps, warm):
    """item-parallel sharding of ONE scene over `devices`, driven from THIS process through stf_group_* (the way a Julia
    host would drive several GPUs): every GPU holds a replica, generates and tests the candidate pairs of its share of the
    sorted records, the GPUs exchange their hit lists themselves over NVLink peer memory (device-side flag barrier), every
    GPU runs the same batchsteps.  One step = the c1 step (restore + reset + round) on every replica.  Timed with CUDA
    events on every GPU's stream (max over GPUs); L2 flushed on every GPU between iterations."""
    from stuffing_b200 import parallel
    W = load_workload(which)
    s = W["mask_side"]
    trees = [S.maskqtree(np.ones((s, s), bool))] + [S.qtree(o, W["canvas"]) for o in W["objs"]]
    for t, r, c in zip(trees, W["rs"], W["cs"]):
        t.shift1 = (int(r), int(c))
    g = parallel.GpuGroup(trees, devices)
    lib = g.lib
    movable = np.arange(2, len(trees) + 1, dtype=np.int32)
    rs = np.ascontiguousarray(W["rs"][1:]); cs = np.ascontiguousarray(W["cs"][1:])
    streams = [torch.cuda.ExternalStream(lib.stf_stream(d.h), device=torch.device("cuda", dv)) for d, dv in zip(g.replicas, g.devices)]
    flush = [torch.empty(512 << 20, dtype=torch.uint8, device=torch.device("cuda", dv)) for dv in g.devices]
    for d in g.replicas:
        lib.stf_set_optimiser(d.h, 2, 0.25, 0.5)

    def step():
        g.setshift(movable, 1, rs, cs)
        g.reset_velocity()
        return g.fit_round(SEED, ITERATION)

    for _ in range(warm):
        nh = step()
    assert nh == W["n_hits"], (nh, W["n_hits"])
    items = sum(d.counters()["items"] + d.counters()["direct"] for d in g.replicas)
    assert items == W["n_items"], (items, W["n_items"])
    tot = 0.0
    for _ in range(steps):
        for st, f in zip(streams, flush):
            with torch.cuda.stream(st):
                f.zero_()
        for dv in g.devices:
            torch.cuda.synchronize(dv)
        ev = []
        for st in streams:
            a = torch.cuda.Event(enable_timing=True); a.record(st); ev.append(a)
        step()
        ms = 0.0
        for st, a in zip(streams, ev):
            b = torch.cuda.Event(enable_timing=True); b.record(st); b.synchronize()
            ms = max(ms, a.elapsed_time(b))
        tot += ms
    out = {"workload": which, "gpus": len(devices), "ms_per_step": tot / steps, "items_per_s": items / (tot / steps * 1e-3), "items_per_step": int(items), "hits_per_step": int(nh),
           "nvlink_bytes_per_step": g.nvlink_bytes, "host_syncs_per_step": len(devices), "exchange": "peer stores over NVLink + device-side flag barrier (stf_group_fit_round), one process drives all GPUs"}
    g.close()
    return out


def run_group(args):
    """`--shard items`: rank 0 drives ALL N GPUs of the box from one process (stf_group_*); the other ranks of a torchrun
    launch only keep the barrier.  Reports the N-GPU group next to the 1-GPU group of the same step (strong scaling)."""
    import torch
    import _pkg
    S = _pkg.load()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = None
    host_group = dist.new_group(backend="gloo") if dist is not None else None
    if rank == 0:
        n = max(world, args.gpus)
        n = min(n, torch.cuda.device_count())
        warm = max(3, args.warmup)
        sampler = ClockSampler(0)
        one = group_bench(S, torch, args.workload, [0], args.steps, warm)
        many = group_bench(S, torch, args.workload, list(range(n)), args.steps, warm) if n > 1 else one
        clocks = sampler.stop()
        peak, peak_src = peaks()
        line = {"metric": "collision_queries_per_sec", "value": many["items_per_s"], "unit": "items/s", "n_gpus": n, "steps": args.steps, "warmup": warm,
                "ms_per_step": many["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "fit_iters_per_sec": 1e3 / many["ms_per_step"], "julia": julia_probe(), "config": {"workload": NAMES[args.workload], "l2": L2_NOTE},
                "details": {"parallelism": f"item-parallel group of {n} GPUs driven by one process", "one_gpu": one, "n_gpu": many, "speedup_vs_one_gpu": one["ms_per_step"] / many["ms_per_step"]},
                "e2e": {"value": many["items_per_s"], "unit": "items/s", "ms_per_step": many["ms_per_step"], "h2d_bytes_per_step": int(12 * 10000 * n), "d2h_bytes_per_step": int(136 * n)},
                "nvlink_bytes_per_step": many["nvlink_bytes_per_step"], "gpu_launches": None, "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src,
                             "note": "latency-bound round; see the default bench line for the per-stage roofline"},
                "cpu_baseline": {"value": None, "unit": "items/s", "cores": 0, "kind": "port", "sample": "see the default bench line"}}
    if dist is not None:
        dist.barrier(group=host_group)  # host-side: no kernel may sit on the other GPUs while rank 0 drives them
        dist.destroy_process_group()
    if line:
        print(json.dumps(line), flush=True)


def run_c3(args):
    """BASELINE configs[2]: the dynamic loop of examples/dynamiccollisions.jl / trainepoch_D! (fit.jl:224-233) on the C1 scene.
    One STEP = 1 % of the trees jittered by +-(1..8) px (shift!), partialcollisions over the moved subset (the labels of the
    previous result + the jittered ones) through the LinkedSpacialQTree, batchsteps! on the result; the labels of the result
    come back to the HOST every step (the reference's control flow needs them: moved := labels of the result).  The state
    evolves, so steps differ; the line reports totals over the K timed steps.  Every step crosses the boundary with host
    buffers: value == e2e."""
    import torch
    import _pkg
    S = _pkg.load()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = load_workload("c1")
    d = workloads.device_scene(S, W, local)
    lib, h, n_all = d.lib, d.h, len(d)
    stream = torch.cuda.ExternalStream(lib.stf_stream(h), device=torch.device("cuda", local))
    lib.stf_set_optimiser(h, 2, 0.25, 0.5)
    tree = S.linked_spacial_qtree(d)
    S.locate(d, None, tree)
    moved = list(range(2, n_all + 1))
    tot = dict(items=0, hits=0, moved=0, ms=0.0, h2d=0, d2h=0)
    nh, nm, moved_buf = C.c_int64(0), C.c_int32(0), np.zeros(n_all, np.int32)
    warm = max(3, args.warmup)
    script = c3_script(n_all, warm + args.steps)
    l0 = 0
    sampler = None
    for step in range(warm + args.steps):
        timed = step >= warm
        if step == warm:
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            sampler = ClockSampler(local) if rank == 0 else None
            l0 = lib.stf_launch_count(h)
        jit, dr, dc = script[step]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        d.shift(jit, 1, dr, dc)
        labels = np.asarray(list(dict.fromkeys(moved + jit.tolist())), np.int32)
        # ONE call: partialcollisions over `labels` + batchsteps! on the device; only the labels of the result come back
        d._ck(lib.stf_fit_round_moved(h, 1, len(labels), labels.ctypes.data, SEED, step, C.byref(nh), moved_buf.ctypes.data, len(moved_buf), C.byref(nm)))
        b.record(stream)
        b.synchronize()
        c = d.counters()
        res = moved_buf[: nm.value]
        moved = res[res != 1].tolist()
        if timed:
            tot["ms"] += a.elapsed_time(b); tot["items"] += c["items"] + c["direct"]; tot["hits"] += nh.value; tot["moved"] += len(labels)
            tot["h2d"] += 4 * len(labels) + 12 * len(jit); tot["d2h"] += 4 * nm.value + 136
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    launches = lib.stf_launch_count(h) - l0
    t = torch.tensor([tot["ms"]], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / args.steps
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    items = tot["items"] / args.steps
    alg = 16.0 * items + 0.5 * 7.4 * items + 40.0 * tot["moved"] / args.steps + 100.0 * tot["hits"] / args.steps
    line = {"metric": "collision_queries_per_sec", "value": world * items / (ms * 1e-3), "unit": "items/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "fit_iters_per_sec": world * 1e3 / ms,
            "julia": julia_probe(), "config": {"workload": NAMES["c3"], "l2": L2_NOTE},
            "details": {"objects": n_all, "items_per_step": int(items), "hits_per_step": int(tot["hits"] / args.steps),
                        "moved_labels_per_step": int(tot["moved"] / args.steps),
                        "parallelism": f"scene replicas x{world}"},
            "e2e": {"value": world * items / (ms * 1e-3), "unit": "items/s", "ms_per_step": ms, "h2d_bytes_per_step": int(tot["h2d"] / args.steps), "d2h_bytes_per_step": int(tot["d2h"] / args.steps)},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole dynamic step", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                         "peak_source": peak_src, "note": "latency-bound loop: per step one shift! call (no sync) and one stf_fit_round_moved call (one sync); the hit list stays on the device"},
            "cpu_baseline": {"value": None, "unit": "items/s", "cores": 0, "kind": "port", "sample": "run `bench.py --impl reference --workload c3`: the same scripted dynamic steps on the C restatement"}}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--shard", default="scenes", choices=["scenes", "items"])
    ap.add_argument("--no-build-roofline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded records (C5 scene split, C1 item-parallel group) of the default line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.shard == "items" and args.workload in ("c1", "c4"):
        run_group(args)
    elif args.workload == "c3":
        run_c3(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
