"""GPU experiment: cycle breakdown of the single-CTA block sort (build with -DSTF_BS_DEBUG)"""
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
out = np.zeros(64, np.int64)
lib.stf_debug_sort_cycles.argtypes = [C.c_void_p]
def show(tag):
    lib.stf_debug_sort_cycles(out.ctypes.data)
    print(tag, "n", out[4], "npass", out[3], "plan", out[1] - out[0], "passes", out[2] - out[1], "last pass: zero+rank", out[9] - out[8], "scan", out[10] - out[9], "scatter", out[11] - out[10], "reload", out[12] - out[11])
hits = S.totalcollisions_spacial(d, unique=False, raw=True)
show("last sort of totalcollisions (hits):")
lib.stf_set_optimiser(h, 2, 0.25, 0.5)
nh = C.c_int64(0)
d._ck(lib.stf_fit_round(h, 0, 0, None, 1, 0, C.byref(nh)))
show("last sort of fit_round (endpoints):")
