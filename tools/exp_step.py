"""GPU experiment: where does k_batchsteps spend its time?  (STF_DEBUG_STEP variations break parity; timing only)"""
import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, _pkg
S = _pkg.load()
W = bench.load_workload()
s = W["mask_side"]
trees = [S.maskqtree(np.ones((s, s), bool))] + [S.qtree(o, W["canvas"]) for o in W["objs"]]
for t, r, c in zip(trees, W["rs"], W["cs"]):
    t.shift1 = (int(r), int(c))
os.environ["STF_DEBUG_STEP"] = sys.argv[1] if len(sys.argv) > 1 else "16"  # read once, at stf_create
d = S.DeviceQTrees(trees); lib, h = d.lib, d.h
n_all = len(d); movable = np.arange(2, n_all + 1, dtype=np.int32)
rs, cs = W["rs"][1:].copy(), W["cs"][1:].copy()
lib.stf_set_optimiser(h, 2, 0.25, 0.5)
hits = S.totalcollisions_spacial(d, unique=False, raw=True)
print("hits", len(hits))
for it in range(3):
    d._ck(lib.stf_set_shifts(h, len(movable), movable.ctypes.data, 1, rs.ctypes.data, cs.ctypes.data, 0))
    d._ck(lib.stf_reset_velocity(h, 0, None))
    d._ck(lib.stf_batchsteps(h, len(hits), hits.ctypes.data, 1, 0))
n = len(hits)
T = np.zeros((n, 4), np.uint64)
lib.stf_debug_step_timeline.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
print("rc", lib.stf_debug_step_timeline(h, n, T.ctypes.data))
seg = np.stack([((T[:, 0] >> np.uint64(12 * i)) & np.uint64(0xfff)).astype(np.int64) * 16 for i in range(5)], 1)
print("cycles state-loads/grad/scalar/update-stores/fence: median", np.median(seg, 0).tolist(), "p90", np.percentile(seg, 90, 0).tolist(), "max", seg.max(0).tolist())
T = T.astype(np.int64); T[:, 0] = T[:, 1]; t0 = T[:, 1].min(); T -= t0
print("claim  min/median/max us", T[:, 0].min() / 1e3, np.median(T[:, 0]) / 1e3, T[:, 0].max() / 1e3)
print("ready  min/median/max us", T[:, 1].min() / 1e3, np.median(T[:, 1]) / 1e3, T[:, 1].max() / 1e3)
print("end    min/median/max us", T[:, 3].min() / 1e3, np.median(T[:, 3]) / 1e3, T[:, 3].max() / 1e3)
print("decide-ready median us", np.median(T[:, 2] - T[:, 1]) / 1e3, "end-decide median us", np.median(T[:, 3] - T[:, 2]) / 1e3, "max", (T[:, 3] - T[:, 2]).max() / 1e3)
# dependency structure
last = {}; pred = np.full((n, 2), -1)
for k in range(n):
    for s_, x in enumerate((int(hits[k]["i1"]), int(hits[k]["i2"]))):
        if x != 1:
            if x in last: pred[k, s_] = last[x]
            last[x] = k
lat = []
for k in range(n):
    ps = [p for p in pred[k] if p >= 0]
    if ps:
        lat.append(T[k, 1] - max(T[p, 3] for p in ps))
lat = np.array(lat)
print("flag latency (ready - pred end): median us", np.median(lat) / 1e3, "p90", np.percentile(lat, 90) / 1e3, "max", lat.max() / 1e3, "n", len(lat))
order = np.argsort(T[:, 3])[-5:]
for k in order:
    print("late step", k, dict(hits=tuple(int(x) for x in hits[k]), claim=T[k, 0] / 1e3, ready=T[k, 1] / 1e3, decided=T[k, 2] / 1e3, end=T[k, 3] / 1e3, pred=pred[k].tolist()))
