"""diagnostic: a crowded scene (every object in ~n steps of one round) against the oracle, round by round"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import rect_scene, oracle, product, device_from_oracle
S = product(); orc = oracle()
for n in [int(x) for x in sys.argv[1:]] or [40, 100, 115]:
    qo = rect_scene(orc, n, 3, sz=60, masksize=(500, 500), place=False)
    for k, t in enumerate(qo[1:]):
        t.setshift(1, 270 + k % 3, 270 + k % 2)
    d = device_from_oracle(S, qo)
    oo = orc.Opt("momentum", 0.25, 0.5, len(qo))
    clock = S.trainer.StepClock(seed=2)
    for it in range(3):
        hits = sorted(orc.totalcollisions_spacial(qo, unique=False))
        nh, rs, cs = S.trainer.fit_round_positions(d, optimiser=S.Momentum(0.25, 0.5), clock=clock)
        orc.batchsteps(qo, hits, oo, seed=2, iteration=it)
        want = [t.getshift() for t in qo]
        got = [(int(a), int(b)) for a, b in zip(rs, cs)]
        bad = [i for i in range(len(want)) if want[i] != got[i]]
        lv = {}
        for h in hits: lv[h[1][0]] = lv.get(h[1][0], 0) + 1
        print(f"n={n} it={it} hits={len(hits)} dev={nh} mismatches={len(bad)} first={bad[:5]} levels={sorted(lv.items())}", flush=True)
        if bad: break
