#!/usr/bin/env python
"""bench.py — collision queries/s and fit iterations/s of the collision-and-fit hot path (BASELINE.json).

Default workload (`--workload c1`, the configuration the metric is quoted on): ONE scene of 10 000 word masks + the
mask on a 4096^2 canvas.  One STEP = one fit iteration from a fixed seeded layout:
    restore the level-1 shifts of the 10 000 objects (setshift!: shift update + pyramid rebuild, subsystem 1)
    -> totalcollisions_spacial(qtrees; unique=false): locate!, sort, candidate pairs, frontier BFS (subsystems 2, 3)
    -> batchsteps!: grad2d, step!/move!, Momentum, one rebuild per moved tree (subsystem 4).
Every step does identical work (same layout, same draws), so steps are comparable.

`value`  : narrow-phase items answered per second, whole job, inputs (the shift arrays) resident in HBM.
`e2e`    : the same through the C-ABI call with HOST buffers: pinned shifts H2D every step, hit count and the final
           shifts D2H every step.
`roofline` : the slowest stage of the step against the measured HBM peak (algorithmic bytes), all other
           stages under `stages`, and under `bandwidth_shaped` the one kernel family of the path that streams HBM — the
           fused pack+build of the 4096-scene batch (C5, 1.5 GB of u8 masks resident in HBM).
Other workloads (`--workload c4`: 100k masks / 16384^2; `c5`: 4096 scenes x 129 trees / 512^2 in one batch) run the same
step through the multi-launch throughput path.
N > 1    : c1 / c4 — scene replicas, every rank runs the step on its own copy ("weak"); c5 — the 4096 scenes are split
           over the ranks ("strong"); `--shard items` — item-parallel sharding of ONE scene with an all-gather per step.
--impl reference : the CPU restatement of the reference algorithm (oracle/, kind "port": Julia is not in the image,
           the reference has no C sources) on the host cores, same step, same metric.
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

SEED, ITERATION = 1, 0
FIXTURES = {"c1": "c1_word_n10000_seed1.npz", "c4": "c4_rect_n100000_seed4.npz"}
NAMES = {"c1": "C1 fit iteration: 10k word masks + mask, 4096^2 canvas",
         "c3": "C3 dynamic steps (dynamiccollisions.jl style): 10k word masks + mask, 4096^2 canvas, partialcollisions over the moved subset",
         "c4": "C4 fit iteration: 100k rect masks + mask, 16384^2 canvas",
         "c5": "C5 fit iteration: batch of 4096 scenes x (mask + 128 word masks), 512^2 canvas"}


# ---------------------------------------------------------------------------------------------- workloads
def load_workload(which="c1"):
    """single-scene workloads: seeded objects + the committed place! layout (tools/make_bench_fixture.py)"""
    fx = np.load(os.path.join(ROOT, "bench_data", FIXTURES[which]))
    n, seed, s = int(fx["n"]), int(fx["seed"]), int(fx["mask_side"])
    rng = np.random.default_rng(seed)
    objs = workloads.rect_objects(n, 30, rng) if which == "c4" else workloads.word_objects(n, rng)
    return dict(n=n, mask_side=s, canvas=int(fx["canvas"]), objs=objs, rs=fx["rs"].astype(np.int32), cs=fx["cs"].astype(np.int32),
                n_items=int(fx["n_items"]), n_hits=int(fx["n_hits"]), visited=int(fx["visited"]))


def scene_batch(S, scenes, per_scene=128, seed=5, side=490, canvas=512):
    """C5: `scenes` = iterable of scene ids; every scene = a side^2 mask + per_scene word masks at uniform positions.
    Returns the packed upload (payload, kh, kw, rs, cs, dflt, is_mask, scene ids 0..len-1, canvas)."""
    pool_rng = np.random.default_rng(seed)
    pool = [S.qtree(o, canvas) for o in workloads.word_objects(2000, pool_rng)]
    mask = S.maskqtree(np.ones((side, side), bool))
    mflat = np.asfortranarray(mask.codes).ravel(order="F")
    pflat = [np.asfortranarray(t.codes).ravel(order="F") for t in pool]
    ph, pw = np.array([t.codes.shape[0] for t in pool]), np.array([t.codes.shape[1] for t in pool])
    off = (canvas - side) // 2
    chunks, kh, kw, rs, cs, dflt, is_mask, scene = [], [], [], [], [], [], [], []
    for k, s in enumerate(scenes):
        rng = np.random.default_rng(50_000 + int(s))
        pick = rng.integers(0, len(pool), per_scene)
        chunks.append(mflat); kh.append(side); kw.append(side); rs.append(mask.shift1[0]); cs.append(mask.shift1[1]); dflt.append(mask.default); is_mask.append(1); scene.append(k)
        for j in pick:
            chunks.append(pflat[j]); kh.append(ph[j]); kw.append(pw[j]); dflt.append(pool[j].default); is_mask.append(0); scene.append(k)
            rs.append(off + int(rng.integers(0, max(1, side - ph[j])))); cs.append(off + int(rng.integers(0, max(1, side - pw[j]))))
    return dict(payload=np.concatenate(chunks), kh=np.array(kh, np.int32), kw=np.array(kw, np.int32), rs=np.array(rs, np.int32), cs=np.array(cs, np.int32),
                dflt=np.array(dflt, np.uint8), is_mask=np.array(is_mask, np.uint8), scene=np.array(scene, np.int32), canvas=canvas)


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
def oracle_scene(orc, W):
    return orc.qtrees(W["objs"], mask=np.ones((W["mask_side"], W["mask_side"]), bool))


def oracle_step(orc, qts, W, cores):
    for t, r, c in zip(qts, W["rs"], W["cs"]):
        t.setshift(1, int(r), int(c))
    opt = orc.Opt("momentum", 0.25, 0.5, len(qts))
    hits, info = orc.totalcollisions_spacial(qts, unique=False, nthreads=cores, raw=True)
    orc.lib().orc_sort_unique(hits.ctypes.data, C.byref(C.c_int64(len(hits))), 0)
    orc.batchsteps(qts, hits, opt, seed=SEED, iteration=ITERATION)
    return info["n_items"], len(hits)


def best_threads(orc, qts, W, cores):
    """the thread count (1 or all cores) that runs the narrow phase faster on this host; shared vCPUs can make
    OpenMP slower than one thread"""
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


def run_reference(args):
    """times the CPU restatement on rank 0's host cores; c5 runs a bounded sample of the batch (scene by scene)"""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import oracle as orc  # the one place bench.py executes oracle/: the CPU baseline
    cores_all = os.cpu_count() or 1
    if args.workload == "c5":
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
            scenes.append((ts, dict(rs=B["rs"][s * per:(s + 1) * per], cs=B["cs"][s * per:(s + 1) * per])))
        cores = 1

        def step():
            items = hits = 0
            for ts, W in scenes:
                a, b = oracle_step(orc, ts, W, cores)
                items += a; hits += b
            return items, hits
        sample = f"{nsc} of the 4096 scenes per step, scene after scene on one core"
    else:
        W = load_workload(args.workload)
        qts = oracle_scene(orc, W)
        cores = best_threads(orc, qts, W, cores_all)

        def step():
            return oracle_step(orc, qts, W, cores)
        sample = "full steps (restore shifts + totalcollisions_spacial + batchsteps) of the C restatement, OpenMP over items"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    items = hits = 0
    for _ in range(args.steps):
        items, hits = step()
    ms = (time.perf_counter() - t0) / args.steps * 1e3
    v = items / (ms * 1e-3)
    line = {"impl": "reference", "metric": "collision_queries_per_sec", "value": v, "unit": "items/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "fit_iters_per_sec": 1e3 / ms,
            "config": {"workload": NAMES[args.workload], "items_per_step": items, "hits_per_step": hits},
            "cpu_baseline": {"value": v, "unit": "items/s", "cores": cores, "kind": "port", "sample": f"{args.steps} x {sample}"},
            "e2e": {"value": v, "unit": "items/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(which):
    """the oracle port timed on the host cores for a bounded sample of the same workload (c1: 3 full steps)"""
    try:
        from oracle import oracle as orc
        if which != "c1":
            return {"value": None, "unit": "items/s", "cores": 0, "kind": "port", "sample": "run `bench.py --impl reference --workload %s`" % which}
        W = load_workload(which)
        qts = oracle_scene(orc, W)
        cores = best_threads(orc, qts, W, os.cpu_count() or 1)
        ts, items = [], 0
        for _ in range(3):
            t0 = time.perf_counter()
            items, _h = oracle_step(orc, qts, W, cores)
            ts.append(time.perf_counter() - t0)
        best = min(ts)
        return {"value": items / best, "unit": "items/s", "cores": cores, "kind": "port",
                "sample": "3 full steps (restore shifts + totalcollisions_spacial + sort + batchsteps) of the same 10k-object scene, best of 3", "ms_per_step": best * 1e3}
    except Exception as e:  # the baseline is reporting, never the product path
        return {"value": None, "unit": "items/s", "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}


# ---------------------------------------------------------------------------------------------- our arm (GPU)
def build_roofline(S, torch, local, peak):
    """the bandwidth-shaped kernels of the path: fused level-1 pack + pyramid build of the 4096-scene C5 batch whose u8
    payload is resident in HBM (1.5 GB >> L2).  Algorithmic bytes: the u8 read + 0.25 B per level-1 node
    written + 5 x 0.25 B per upper node (4 children read, 1 written; the fused kernels never re-read them from HBM)."""
    B = scene_batch(S, range(4096))
    dev = torch.from_numpy(np.concatenate([B["payload"], np.zeros(64, np.uint8)])).cuda(local)
    lib = S._lib.load()
    h = C.c_void_p()
    assert lib.stf_create(local, C.byref(h)) == 0
    lib.stf_profile_enable(h, 1)
    L = int(np.log2(B["canvas"])) + 1
    a, b, nodes = B["kh"].astype(np.int64), B["kw"].astype(np.int64), 0
    for l in range(1, L + 1):
        nodes += int((a * b).sum()) * (1 if l == 1 else 5)
        a, b = a // 2 + 1, b // 2 + 1
    alg = float(B["payload"].size) + 0.25 * nodes
    ms = np.zeros(9); calls = np.zeros(9, np.int64)
    best = None
    for _ in range(4):
        rc = lib.stf_upload_packed(h, len(B["kh"]), C.c_void_p(dev.data_ptr()), 1, B["kh"].ctypes.data, B["kw"].ctypes.data, B["rs"].ctypes.data, B["cs"].ctypes.data,
                                   B["dflt"].ctypes.data, B["canvas"], B["is_mask"].ctypes.data, B["scene"].ctypes.data)
        assert rc == 0, lib.stf_last_error(h)
        lib.stf_profile_read(h, ms.ctypes.data, calls.ctypes.data, 1)
        best = ms[8] if best is None else min(best, ms[8])
    lib.stf_destroy(h)
    ach = alg / (best * 1e-3) / 1e9
    return {"kernel": "k_pack_build_small + k_pack_build_mid<128|512> (upload build of the 4096 C5 scenes, 528k trees)", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "ms": float(best), "algorithmic_bytes": int(alg), "payload_bytes": int(B["payload"].size), "input": "u8 payload resident in HBM, larger than L2"}


def run_ours(args):
    import torch
    import _pkg
    S = _pkg.load()
    from stuffing_b200 import parallel
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
    shard_items = args.shard == "items" and which != "c5"

    # ---- the scene(s): upload + full pyramid build on the GPU
    t_up = time.perf_counter()
    if which == "c5":
        my = parallel.shard_range(4096, rank, world)
        B = scene_batch(S, my)
        d = S.DeviceQTrees.from_packed(B["payload"], B["kh"], B["kw"], B["rs"], B["cs"], B["dflt"], B["canvas"], B["is_mask"], B["scene"], device=local)
        rs0, cs0 = B["rs"], B["cs"]
        movable = np.flatnonzero(B["is_mask"] == 0).astype(np.int32) + 1
        objs_shapes = list(zip(B["kh"][movable - 1].tolist(), B["kw"][movable - 1].tolist()))
        scaling, par = "strong", f"scene-parallel: {len(my)} of 4096 scenes per rank x{world}"
    else:
        W = load_workload(which)
        s = W["mask_side"]
        trees = [S.maskqtree(np.ones((s, s), bool))] + [S.qtree(o, W["canvas"]) for o in W["objs"]]
        for t, r, c in zip(trees, W["rs"], W["cs"]):
            t.shift1 = (int(r), int(c))
        d = S.DeviceQTrees(trees, device=local)
        rs0, cs0 = W["rs"], W["cs"]
        movable = np.arange(2, len(trees) + 1, dtype=np.int32)  # the mask (label 1) never moves: restore objects 2..n only
        objs_shapes = [o.shape for o in W["objs"]]
        scaling = "strong" if shard_items else "weak"
        par = f"item-parallel shards of one scene x{world} (all-gather of hit records per step)" if shard_items else f"scene replicas x{world}"
    t_up = time.perf_counter() - t_up
    lib, h = d.lib, d.h
    n_all, nm = len(d), len(movable)
    stream = torch.cuda.ExternalStream(lib.stf_stream(h), device=torch.device("cuda", local))
    rs_h = torch.from_numpy(np.ascontiguousarray(rs0[movable - 1])).pin_memory(); cs_h = torch.from_numpy(np.ascontiguousarray(cs0[movable - 1])).pin_memory()
    rs_d, cs_d = rs_h.cuda(), cs_h.cuda()
    movable_d = torch.from_numpy(movable).cuda()  # resident runs: the label list lives in HBM like the positions
    out_rs = torch.empty(n_all, dtype=torch.int32).pin_memory(); out_cs = torch.empty(n_all, dtype=torch.int32).pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    lib.stf_set_optimiser(h, 2, 0.25, 0.5)
    nh = C.c_int64(0)
    sharded = parallel.ShardedFit(d, rank, world, dist) if shard_items else None

    def step(resident):
        if resident:
            d._ck(lib.stf_set_shifts(h, nm, movable_d.data_ptr(), 1, rs_d.data_ptr(), cs_d.data_ptr(), 0 | 2 | 4))
        else:
            d._ck(lib.stf_set_shifts(h, nm, movable.ctypes.data, 1, rs_h.data_ptr(), cs_h.data_ptr(), 0))
        d._ck(lib.stf_reset_velocity(h, 0, None))
        if sharded is not None:
            nh.value = sharded.round(SEED, ITERATION)
            if not resident:
                d._ck(lib.stf_get_shifts(h, n_all, None, 1, out_rs.data_ptr(), out_cs.data_ptr()))
        elif resident:
            d._ck(lib.stf_fit_round(h, 0, 0, None, SEED, ITERATION, C.byref(nh)))
        else:  # the user-facing call: the round and the positions after it
            d._ck(lib.stf_fit_round_positions(h, 0, 0, None, SEED, ITERATION, C.byref(nh), 1, out_rs.data_ptr(), out_cs.data_ptr()))
        return nh.value

    def timed(resident, steps):
        tot = 0.0
        for _ in range(steps):
            with torch.cuda.stream(stream):
                flush.zero_()  # L2 flush between timed iterations, outside the event pair
            if sharded is not None:
                torch.cuda.synchronize()
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

    warm = max(3, args.warmup)
    for _ in range(warm):
        step(True); step(False)
    c = d.counters()
    items, hits = c["items"] + c["direct"], c["hits"]

    sampler = ClockSampler(local) if rank == 0 else None
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
    c2 = d.counters()
    if sharded is None:
        assert c2["items"] + c2["direct"] == items and c2["hits"] == hits, "steps must do identical work"

    t = torch.tensor([ms_res, ms_e2e, float(items), float(hits), float(c["visited"])], dtype=torch.float64, device="cuda")
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
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline (SURVEY §8d algorithmic bytes): every stage, the slowest one on top
    per_step = stage_ms[:8] / args.steps
    names = S._lib.STAGES
    L = d.nlevels
    nrec = c["records"]
    # the restore rebuilds only the trees the previous step moved (stf_set_shifts leaves a tree that already sits at the
    # requested shift alone): count those, from the positions the last end-to-end step read back
    mv = movable - 1
    moved_sel = (out_rs.numpy()[mv] != np.asarray(rs0)[mv]) | (out_cs.numpy()[mv] != np.asarray(cs0)[mv])
    up_nodes = float(sum(5 * 0.25 * ((((kh - 2) >> (l - 1)) + 2) * (((kw - 2) >> (l - 1)) + 2))
                         for (kh, kw), m in zip(objs_shapes, moved_sel.tolist()) if m for l in range(2, L + 1)))
    alg = {  # algorithmic bytes per step for each stage (this rank)
        "shift_rebuild": up_nodes + 8.0 * nm,
        "locate": 16.0 * n_all + 0.25 * 4 * 6 * n_all,
        "sort": 2 * 12.0 * nrec,
        "emit": 16.0 * items,
        "narrow": 16.0 * items + 0.5 * c["visited"] + 16.0 * hits,
        "result": 2 * 16.0 * hits,
        "step_prep": 2 * 12.0 * 2 * hits,
        "step": hits * (2 * 8 * 0.25 + 2 * (8 + 16)) + 1.25 * 60 * c["moved"],
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
    if not args.no_build_roofline:
        try:
            roof["bandwidth_shaped"] = build_roofline(S, torch, local, peak)
            roof["bandwidth_shaped"]["traffic"] = traffic.get("build")
        except Exception as e:  # reporting only
            roof["bandwidth_shaped"] = {"error": str(e)}

    cpu = cpu_baseline(which)
    scenes_job = 4096 if which == "c5" else (1 if shard_items else world)
    line = {"metric": "collision_queries_per_sec", "value": items_job / (ms_res * 1e-3), "unit": "items/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_res, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "fit_iters_per_sec": scenes_job * 1e3 / ms_res,
            "config": {"workload": NAMES[which] + " (restore shifts + rebuild, totalcollisions_spacial, batchsteps)",
                       "objects": n_all, "levels": L, "items_per_step": int(items_job), "hits_per_step": int(hits_job),
                       "trees_rebuilt_by_restore_per_step": int(moved_sel.sum()),
                       "visited_nodes_per_step": int(float(t[4])) if which == "c5" or shard_items else c["visited"] * world,
                       "l2": "flushed between timed iterations (512 MiB write)", "parallelism": par, "upload_build_s": round(t_up, 3)},
            "e2e": {"value": items_job / (ms_e2e * 1e-3), "unit": "items/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(12 * nm), "d2h_bytes_per_step": int(8 * n_all + 8 + 96)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    if sharded is not None:
        line["nvlink_bytes_per_step"] = int(sharded.bytes_exchanged / max(1, 2 * (args.steps + warm)))
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_c3(args):
    """BASELINE configs[2]: the dynamic loop of examples/dynamiccollisions.jl / trainepoch_D! (fit.jl:224-233) on the C1 scene.
    One STEP = 1 % of the trees jittered by +-(1..8) px (shift!), partialcollisions over the moved subset (the labels of the
    previous result + the jittered ones) through the LinkedSpacialQTree, the result returned to the HOST (the reference's
    control flow needs it: moved := labels of the result), batchsteps! on it.  The state evolves, so steps differ; the line
    reports totals over the K timed steps.  Every step crosses the boundary with host buffers: value == e2e."""
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
    s_ = W["mask_side"]
    trees = [S.maskqtree(np.ones((s_, s_), bool))] + [S.qtree(o, W["canvas"]) for o in W["objs"]]
    for t, r, c in zip(trees, W["rs"], W["cs"]):
        t.shift1 = (int(r), int(c))
    d = S.DeviceQTrees(trees, device=local)
    lib, h, n_all = d.lib, d.h, len(trees)
    stream = torch.cuda.ExternalStream(lib.stf_stream(h), device=torch.device("cuda", local))
    lib.stf_set_optimiser(h, 2, 0.25, 0.5)
    tree = S.linked_spacial_qtree(d)
    S.locate(d, None, tree)
    rng = np.random.default_rng(3)
    moved = list(range(2, n_all + 1))
    tot = dict(items=0, hits=0, moved=0, ms=0.0, h2d=0, d2h=0)
    warm = max(3, args.warmup)
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
        jit = rng.choice(np.arange(2, n_all + 1), size=n_all // 100, replace=False).astype(np.int32)
        dr = (rng.integers(1, 9, len(jit)) * rng.choice([-1, 1], len(jit))).astype(np.int32)
        dc = (rng.integers(1, 9, len(jit)) * rng.choice([-1, 1], len(jit))).astype(np.int32)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        d.shift(jit, 1, dr, dc)
        labels = list(dict.fromkeys(moved + jit.tolist()))
        colist = S.partialcollisions(d, tree, labels, unique=False, raw=True)
        c = d.counters()
        d._ck(lib.stf_batchsteps(h, len(colist), colist.ctypes.data, SEED, step))
        b.record(stream)
        b.synchronize()
        moved = sorted({int(x) for r_ in colist for x in (r_["i1"], r_["i2"]) if x != 1})
        if timed:
            tot["ms"] += a.elapsed_time(b); tot["items"] += c["items"] + c["direct"]; tot["hits"] += len(colist); tot["moved"] += len(labels)
            tot["h2d"] += 4 * len(labels) + 12 * len(jit) + 20 * len(colist); tot["d2h"] += 20 * len(colist) + 128
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
            "config": {"workload": NAMES["c3"], "objects": n_all, "items_per_step": int(items), "hits_per_step": int(tot["hits"] / args.steps),
                       "moved_labels_per_step": int(tot["moved"] / args.steps), "l2": "not flushed: the state evolves from step to step (dynamic loop); the working set (5 MB) is L2-resident by design",
                       "parallelism": f"scene replicas x{world}"},
            "e2e": {"value": world * items / (ms * 1e-3), "unit": "items/s", "ms_per_step": ms, "h2d_bytes_per_step": int(tot["h2d"] / args.steps), "d2h_bytes_per_step": int(tot["d2h"] / args.steps)},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole dynamic step", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                         "peak_source": peak_src, "note": "latency-bound host-driven loop: two C-ABI calls with host lists per step"},
            "cpu_baseline": {"value": None, "unit": "items/s", "cores": 0, "kind": "port", "sample": "see the c1 line: the same scene, one full round"}}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1", choices=["c1", "c3", "c4", "c5"])
    ap.add_argument("--shard", default="scenes", choices=["scenes", "items"])
    ap.add_argument("--no-build-roofline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload == "c3":
            raise SystemExit("bench.py: the reference arm covers c1, c4 and c5")
        run_reference(args)
    elif args.workload == "c3":
        run_c3(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
