import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, _pkg
S = _pkg.load()
W = bench.load_workload()
s = W["mask_side"]
trees = [S.maskqtree(np.ones((s, s), bool))] + [S.qtree(o, W["canvas"]) for o in W["objs"]]
for t, r, c in zip(trees, W["rs"], W["cs"]):
    t.shift1 = (int(r), int(c))
d = S.DeviceQTrees(trees); lib, h = d.lib, d.h
out = np.zeros(16, np.int64)
lib.stf_debug_sort_cycles.argtypes = [C.c_void_p]
S.locate(d)
lib.stf_debug_sort_cycles(out.ctypes.data); print("records sort: after compaction, windows, passes, unpack (cycles):", out[:4].tolist(), "n,npass,packed", out[8:11].tolist(), "last pass hist/scan/scatter", out[4:7].tolist())
hits = S.totalcollisions_spacial(d, unique=False, raw=True)
lib.stf_debug_sort_cycles(out.ctypes.data); print("hits sort   :", out[:4].tolist(), "n,npass,packed", out[8:11].tolist(), "last pass hist/scan/scatter", out[4:7].tolist())
lib.stf_set_optimiser(h, 2, 0.25, 0.5)
d._ck(lib.stf_batchsteps(h, len(hits), hits.ctypes.data, 1, 0))
lib.stf_debug_sort_cycles(out.ctypes.data); print("endpoint sort:", out[:4].tolist(), "n,npass,packed", out[8:11].tolist(), "last pass hist/scan/scatter", out[4:7].tolist())
