"""Generates the layout fixtures bench.py and the full-size tests use: the level-1 shifts that the
reference's place! (restated in the oracle, qtree_functions.jl:372-402,503-543) assigns to the seeded
synthetic objects of workloads.py.  Run once in the build container; the .npz files are committed.

    python tools/make_bench_fixture.py
"""
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


if __name__ == "__main__":
    make("c1_word_n10000_seed1", 10000, 1, "word")
    make("c1_word_n1000_seed1", 1000, 1, "word")
