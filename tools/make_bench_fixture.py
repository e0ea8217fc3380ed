"""Generates the layout fixtures bench.py and the full-size tests use: the level-1 shifts that the
reference's place! (restated in the oracle, qtree_functions.jl:372-402,503-543) assigns to the seeded
synthetic objects of workloads.py.  Run once in the build container; the .npz files are committed.

    python tools/make_bench_fixture.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def make(name, n, seed, kind):
    rng = np.random.default_rng(seed)
    s = int(round(35 * np.sqrt(n)))  # examples/collision_benchmark.jl:23
    objs = workloads.word_objects(n, rng) if kind == "word" else workloads.rect_objects(n, 30, rng)
    qts = orc.qtrees(objs, mask=np.ones((s, s), bool))
    orc.place(qts, seed)
    rs = np.array([t.getshift()[0] for t in qts], np.int32)
    cs = np.array([t.getshift()[1] for t in qts], np.int32)
    hits, det = orc.totalcollisions_spacial(qts, unique=False, details=True)
    path = os.path.join(ROOT, "bench_data", name + ".npz")
    np.savez_compressed(path, rs=rs, cs=cs, n=n, seed=seed, mask_side=s, canvas=qts[0].info(1)[4],
                        n_items=len(det["items"]), n_hits=len(hits), visited=det["visited"])
    print(name, "n", n, "canvas", qts[0].info(1)[4], "items", len(det["items"]), "hits", len(hits), "visited", det["visited"])


def make_c2(name="c2_rect_n1000_seed2", n=1000, seed=2, rounds=50):
    """BASELINE configs[1] (examples/packing.jl fit loop): 1 000 G-rect objects sz = 36 on a 1000^2 mask
    (canvas 1024, mask = label 1), place!, then `rounds` fixed trainer-E-style rounds: totalcollisions(qtrees; unique=false)
    -> batchsteps! with Momentum(1/4, 1/2), canonical draws keyed (seed, round).  The per-round counts and the final
    positions are the golden vector of tests/test_gpu_parity.py::test_c2_* and bench.py --workload c2."""
    rng = np.random.default_rng(seed)
    objs = workloads.rect_objects(n, 36, rng)
    qts = orc.TreeList(orc.qtrees(objs, mask=np.ones((1000, 1000), bool)))
    orc.place(qts, seed)
    rs, cs = orc.getshifts(qts)
    opt = orc.Opt("momentum", 0.25, 0.5, len(qts))
    items, hits, visited = [], [], []
    for k in range(rounds):
        h, info = orc.totalcollisions_spacial(qts, unique=False, raw=True)
        h = h.copy()
        orc.lib().orc_sort_unique(h.ctypes.data, C.byref(C.c_int64(len(h))), 0)
        items.append(info["n_items"] + info["n_direct"]); hits.append(len(h)); visited.append(info["visited"])
        orc.batchsteps(qts, h, opt, seed=seed, iteration=k)
    frs, fcs = orc.getshifts(qts)
    path = os.path.join(ROOT, "bench_data", name + ".npz")
    np.savez_compressed(path, rs=rs, cs=cs, n=n, seed=seed, mask_side=1000, canvas=qts[0].info(1)[4], rounds=rounds,
                        round_items=np.array(items, np.int64), round_hits=np.array(hits, np.int64), round_visited=np.array(visited, np.int64),
                        final_rs=frs, final_cs=fcs, n_items=items[0], n_hits=hits[0], visited=visited[0])
    print(name, "canvas", qts[0].info(1)[4], "items", items[:5], "...", sum(items), "hits", hits[:8], "...", hits[-3:])


if __name__ == "__main__":
    if "c2" in sys.argv:
        make_c2()
        sys.exit(0)
    make("c1_word_n10000_seed1", 10000, 1, "word")
    make("c1_word_n1000_seed1", 1000, 1, "word")
