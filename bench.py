#!/usr/bin/env python
"""bench.py — collision queries/s and fit iterations/s at 10k masks on a 4096^2 canvas (BASELINE.json).

One STEP = one fit iteration of the hot path from a fixed seeded layout:
    restore the 10 001 level-1 shifts (setshift! of every tree: shift update + pyramid rebuild, subsystem 1)
    -> totalcollisions_spacial(qtrees; unique=false): locate!, radix sort, candidate pairs, frontier BFS (2, 3)
    -> batchsteps!: grad2d, step!/move!, Momentum, rebuild of the moved trees (4).
Every step does identical work (same layout, same draws), so steps are comparable.

`value`  : narrow-phase items answered per second, whole job, inputs (the shift arrays) resident in HBM.
`e2e`    : the same through the C-ABI call with HOST buffers: pinned shifts H2D every step, hit count and
           the final shifts D2H every step.
N > 1    : scene-parallel replicas (the path shards by scene; no data-path collective): every rank runs the
           step on its own copy of the scene, value = N * items / max-over-ranks time ("weak").
--impl reference : the CPU restatement of the reference algorithm (oracle/, kind "port": Julia is not in the
           image, the reference has no C sources) on the host cores, same step, same metric.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402

FIXTURE = os.path.join(ROOT, "bench_data", "c1_word_n10000_seed1.npz")
SEED, ITERATION = 1, 0


def load_workload():
    fx = np.load(FIXTURE)
    n, seed, s = int(fx["n"]), int(fx["seed"]), int(fx["mask_side"])
    objs = workloads.word_objects(n, np.random.default_rng(seed))
    return dict(n=n, mask_side=s, canvas=int(fx["canvas"]), objs=objs, rs=fx["rs"].astype(np.int32), cs=fx["cs"].astype(np.int32),
                n_items=int(fx["n_items"]), n_hits=int(fx["n_hits"]), visited=int(fx["visited"]))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.rows, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
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


# ---------------------------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args, W):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc  # the one place bench.py executes oracle/: the CPU baseline
    cores = os.cpu_count() or 1
    qts = orc.qtrees(W["objs"], mask=np.ones((W["mask_side"], W["mask_side"]), bool))
    n_all = len(qts)
    opt = None
    cores = best_threads(orc, qts, W, cores)

    def step():
        nonlocal opt
        for t, r, c in zip(qts, W["rs"], W["cs"]):
            t.setshift(1, int(r), int(c))
        opt = orc.Opt("momentum", 0.25, 0.5, n_all)
        hits, info = orc.totalcollisions_spacial(qts, unique=False, nthreads=cores, raw=True)
        orc.lib().orc_sort_unique(hits.ctypes.data, C.byref(C.c_int64(len(hits))), 0)
        orc.batchsteps(qts, hits, opt, seed=SEED, iteration=ITERATION)
        return info["n_items"], len(hits)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    items = hits = 0
    for _ in range(args.steps):
        items, hits = step()
    dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    v = items / (ms * 1e-3)
    line = {"impl": "reference", "metric": "collision_queries_per_sec", "value": v, "unit": "items/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "fit_iters_per_sec": 1e3 / ms,
            "config": {"workload": "C1 fit iteration: 10k word masks + mask, 4096^2 canvas", "objects": n_all, "items_per_step": items, "hits_per_step": hits},
            "cpu_baseline": {"value": v, "unit": "items/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} full steps (restore shifts + totalcollisions_spacial + batchsteps) of the C restatement, OpenMP over items"},
            "e2e": {"value": v, "unit": "items/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- our arm (GPU)
def run_ours(args, W):
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

    # ---- scene: mask + 10k word objects at the fixture layout (upload + full pyramid build on the GPU)
    s = W["mask_side"]
    trees = [S.maskqtree(np.ones((s, s), bool))] + [S.qtree(o, W["canvas"]) for o in W["objs"]]
    for t, r, c in zip(trees, W["rs"], W["cs"]):
        t.shift1 = (int(r), int(c))
    t_up = time.perf_counter()
    d = S.DeviceQTrees(trees, device=local)
    t_up = time.perf_counter() - t_up
    lib, h = d.lib, d.h
    n_all = len(d)
    stream = torch.cuda.ExternalStream(lib.stf_stream(h), device=torch.device("cuda", local))
    rs_h = torch.from_numpy(W["rs"].copy()).pin_memory(); cs_h = torch.from_numpy(W["cs"].copy()).pin_memory()
    rs_d, cs_d = rs_h.cuda(), cs_h.cuda()
    out_rs = torch.empty(n_all, dtype=torch.int32).pin_memory(); out_cs = torch.empty(n_all, dtype=torch.int32).pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    lib.stf_set_optimiser(h, 2, 0.25, 0.5)
    nh = C.c_int64(0)

    movable = np.arange(2, n_all + 1, dtype=np.int32)  # the mask (label 1) never moves: restore objects 2..n only
    nm = len(movable)

    def step(resident):
        if resident:
            d._ck(lib.stf_set_shifts(h, nm, movable.ctypes.data, 1, rs_d.data_ptr() + 4, cs_d.data_ptr() + 4, 0 | 2))
        else:
            d._ck(lib.stf_set_shifts(h, nm, movable.ctypes.data, 1, rs_h.data_ptr() + 4, cs_h.data_ptr() + 4, 0))
        d._ck(lib.stf_reset_velocity(h, 0, None))
        d._ck(lib.stf_fit_round(h, 0, 0, None, SEED, ITERATION, C.byref(nh)))
        if not resident:
            d._ck(lib.stf_get_shifts(h, n_all, None, 1, out_rs.data_ptr(), out_cs.data_ptr()))
        return nh.value

    def timed(resident, steps):
        tot = 0.0
        for _ in range(steps):
            with torch.cuda.stream(stream):
                flush.zero_()  # L2 flush between timed iterations, outside the event pair
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            step(resident)
            b.record(stream)
            b.synchronize()
            tot += a.elapsed_time(b)
        return tot

    def bracket():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step(True); step(False)
    c = d.counters()
    items, hits = c["items"] + c["direct"], c["hits"]

    sampler = ClockSampler(local) if rank == 0 else None
    lib.stf_profile_enable(h, 1)
    lib.stf_profile_read(h, None, None, 1)
    l0 = lib.stf_launch_count(h)
    bracket()
    ms_res = timed(True, args.steps)
    bracket()
    launches = lib.stf_launch_count(h) - l0
    stage_ms = np.zeros(9); stage_calls = np.zeros(9, np.int64)
    lib.stf_profile_read(h, stage_ms.ctypes.data, stage_calls.ctypes.data, 1)
    lib.stf_profile_enable(h, 0)
    bracket()
    ms_e2e = timed(False, args.steps)
    bracket()
    clocks = sampler.stop() if sampler else None
    c2 = d.counters()
    assert c2["items"] + c2["direct"] == items and c2["hits"] == hits, "steps must do identical work"

    t = torch.tensor([ms_res, ms_e2e], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_res, ms_e2e = (float(x) / args.steps for x in t.tolist())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant stage (SURVEY §8d algorithmic bytes)
    per_step = stage_ms / args.steps
    names = S._lib.STAGES
    L = d.nlevels
    nrec = c["records"]
    alg = {  # algorithmic bytes per step for each stage
        "shift_rebuild": float(sum(5 * 0.25 * ((((o.shape[0] - 2) >> (l - 1)) + 2) * (((o.shape[1] - 2) >> (l - 1)) + 2)) for o in W["objs"] for l in range(2, L + 1))) + 8.0 * n_all,
        "locate": 16.0 * n_all + 0.25 * 4 * 6 * n_all,
        "sort": 2 * 12.0 * nrec,
        "emit": 16.0 * items,
        "narrow": 16.0 * items + 0.5 * c["visited"] + 16.0 * hits,
        "result": 2 * 16.0 * hits,
        "step_prep": 2 * 12.0 * 2 * hits,
        "step": hits * (2 * 8 * 0.25 + 2 * (8 + 16)),
    }
    dom = int(np.argmax(per_step))
    peak, peak_src = peaks()
    dom_ms = per_step[dom] / max(1, stage_calls[dom] / args.steps)
    achieved = alg[names[dom]] / max(1, stage_calls[dom] / args.steps) / (dom_ms * 1e-3) / 1e9
    total_alg = sum(alg.values())
    roof = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "peak_source": peak_src, "stage_ms_per_step": {names[i]: round(float(per_step[i]), 4) for i in range(8)},
            "algorithmic_bytes_per_step": {k: int(v) for k, v in alg.items()},
            "whole_step": {"achieved": total_alg / (ms_res * 1e-3) / 1e9, "frac": total_alg / (ms_res * 1e-3) / 1e9 / peak}}

    # ---- CPU baseline beside it: a bounded sample of the same workload on the host cores
    cpu = cpu_baseline(W, items)
    line = {"metric": "collision_queries_per_sec", "value": world * items / (ms_res * 1e-3), "unit": "items/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "fit_iters_per_sec": world * 1e3 / ms_res,
            "config": {"workload": "C1 fit iteration: 10k word masks + mask, 4096^2 canvas (restore shifts + rebuild, totalcollisions_spacial, batchsteps)",
                       "objects": n_all, "levels": L, "items_per_step": items, "hits_per_step": hits, "visited_nodes_per_step": c["visited"],
                       "l2": "flushed between timed iterations (512 MiB write)", "parallelism": f"scene replicas x{world}", "upload_build_s": round(t_up, 3)},
            "e2e": {"value": world * items / (ms_e2e * 1e-3), "unit": "items/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(12 * (n_all - 1)), "d2h_bytes_per_step": int(8 * n_all + 8 + 64)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def best_threads(orc, qts, W, cores):
    """the thread count (1 or all cores) that runs the narrow phase faster on this host; shared vCPUs can
    make OpenMP slower than one thread"""
    for t, r, c in zip(qts, W["rs"], W["cs"]):
        t.setshift(1, int(r), int(c))
    best, best_t = 1, None
    for nt in sorted({1, cores}):
        orc.totalcollisions_spacial(qts, unique=False, nthreads=nt, raw=True)
        t0 = time.perf_counter()
        orc.totalcollisions_spacial(qts, unique=False, nthreads=nt, raw=True)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = nt, dt
    return best


def cpu_baseline(W, items):
    """the oracle port timed on the host cores for a bounded sample: 3 full steps of the same workload"""
    try:
        from oracle import oracle as orc
        cores = os.cpu_count() or 1
        qts = orc.qtrees(W["objs"], mask=np.ones((W["mask_side"], W["mask_side"]), bool))
        n_all = len(qts)
        cores = best_threads(orc, qts, W, cores)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            for t, r, c in zip(qts, W["rs"], W["cs"]):
                t.setshift(1, int(r), int(c))
            hits, info = orc.totalcollisions_spacial(qts, unique=False, nthreads=cores, raw=True)
            orc.batchsteps(qts, hits, orc.Opt("momentum", 0.25, 0.5, n_all), seed=SEED, iteration=ITERATION)
            ts.append(time.perf_counter() - t0)
        best = min(ts)
        return {"value": info["n_items"] / best, "unit": "items/s", "cores": cores, "kind": "port",
                "sample": "3 full steps (restore shifts + totalcollisions_spacial + batchsteps) of the same 10k-object scene, best of 3", "ms_per_step": best * 1e3}
    except Exception as e:  # the baseline is reporting, never the product path
        return {"value": None, "unit": "items/s", "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    W = load_workload()
    if args.impl == "reference":
        run_reference(args, W)
    else:
        run_ours(args, W)


if __name__ == "__main__":
    main()
