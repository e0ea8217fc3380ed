"""GPU experiment: throughput of the on-device pyramid build (u8 payload resident in HBM -> packed pyramids)"""
import ctypes as C, os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, _pkg, workloads
S = _pkg.load()
lib = S._lib.load()

def scene_batch(n_scenes, per_scene, seed, side=490, canvas=512):
    rng = np.random.default_rng(seed)
    pool = [S.qtree(o, canvas) for o in workloads.word_objects(2000, rng)]
    mask = S.maskqtree(np.ones((side, side), bool))
    chunks, kh, kw, dflt, is_mask, scene = [], [], [], [], [], []
    mflat = np.asfortranarray(mask.codes).ravel(order="F")
    pflat = [np.asfortranarray(t.codes).ravel(order="F") for t in pool]
    for s in range(n_scenes):
        chunks.append(mflat); kh.append(mask.codes.shape[0]); kw.append(mask.codes.shape[1]); dflt.append(mask.default); is_mask.append(1); scene.append(s)
        for k in rng.integers(0, len(pool), per_scene):
            t = pool[int(k)]
            chunks.append(pflat[int(k)]); kh.append(t.codes.shape[0]); kw.append(t.codes.shape[1]); dflt.append(t.default); is_mask.append(0); scene.append(s)
    return np.concatenate(chunks), np.array(kh, np.int32), np.array(kw, np.int32), np.array(dflt, np.uint8), np.array(is_mask, np.uint8), np.array(scene, np.int32), canvas

def run(name, payload, kh, kw, dflt, is_mask, scene, canvas, reps=5):
    n = len(kh)
    dev = torch.from_numpy(np.concatenate([payload, np.zeros(64, np.uint8)])).cuda()
    h = C.c_void_p(); assert lib.stf_create(0, C.byref(h)) == 0
    lib.stf_profile_enable(h, 1)
    ms = np.zeros(9); calls = np.zeros(9, np.int64)
    L = int(np.log2(canvas)) + 1
    up = (payload.size)                       # u8 read
    nodes = 0
    a, b = kh.astype(np.int64).copy(), kw.astype(np.int64).copy()
    for l in range(1, L + 1):
        nodes_l = int((a * b).sum())
        nodes += nodes_l if l == 1 else 5 * nodes_l  # level 1: write; level l>1: read 4 children + write 1
        a = a // 2 + 1; b = b // 2 + 1
    alg = up + 0.25 * nodes
    for r in range(reps):
        t0 = time.perf_counter()
        rc = lib.stf_upload_packed(h, n, C.c_void_p(dev.data_ptr()), 1, kh.ctypes.data, kw.ctypes.data, None, None, dflt.ctypes.data, canvas, is_mask.ctypes.data, scene.ctypes.data)
        assert rc == 0, lib.stf_last_error(h)
        wall = time.perf_counter() - t0
        lib.stf_profile_read(h, ms.ctypes.data, calls.ctypes.data, 1)
        k4 = np.zeros(4); lib.stf_profile_build_kernels(h, k4.ctypes.data)
        print("   per kernel (us): small %.1f  mid128 %.1f  mid512 %.1f  mask path %.1f" % tuple(k4 * 1e3), flush=True)
        print(f"{name}: n={n} payload={up/1e6:.1f} MB algorithmic={alg/1e6:.1f} MB  build kernels {ms[8]*1e3:.1f} us -> {alg/ms[8]/1e6:.1f} GB/s ({alg/ms[8]/1e6/6534.8*100:.1f}% of measured HBM peak); whole call {wall*1e3:.1f} ms", flush=True)
    lib.stf_destroy(h)

if __name__ == "__main__":
    ns = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    run("C5 batch", *scene_batch(ns, 128, 5))
    # one big mask (C1: 3500^2) + 10k words
    rng = np.random.default_rng(1)
    objs = [S.qtree(o, 4096) for o in workloads.word_objects(10000, rng)]
    mask = S.maskqtree(np.ones((3500, 3500), bool))
    ts = [mask] + objs
    payload = np.concatenate([np.asfortranarray(t.codes).ravel(order="F") for t in ts])
    run("C1 scene", payload, np.array([t.codes.shape[0] for t in ts], np.int32), np.array([t.codes.shape[1] for t in ts], np.int32),
        np.array([t.default for t in ts], np.uint8), np.array([1] + [0] * len(objs), np.uint8), np.zeros(len(ts), np.int32), 4096)
