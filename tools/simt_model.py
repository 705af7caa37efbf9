#!/usr/bin/env python
"""Offline SIMT-efficiency model of the batched kernel from the oracle's attempt
logs: 32 instances of one IBIS corner = one warp.  Compares (A) the round-1
schedule (a warp claims 32 instances, attempts in lock-step, Newton loop waits
for the slowest lane), (B) the same with per-lane claiming (no tail), (C) one
LU solve per loop trip for every lane (flattened state machine).
usage: simt_model.py cfg4|cfg5 [c_pro c_lu c_epi]"""
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def one(args):
    cfg, M, i = args
    import backends
    import workloads
    w = workloads.make(cfg, M)
    nl, s = w.instance_netlist(i)
    o = backends.OracleSim(nl, s)
    o.log_attempts()
    o.tran(w.tstep, w.tstop)
    t, f = o.attempt_log()
    return f >> 8, (f >> 4) & 15


if __name__ == "__main__":
    cfg = sys.argv[1]
    cpro, clu, cepi = [float(x) for x in sys.argv[2:5]] if len(sys.argv) > 4 else (1.0, 0.6, 0.6)
    M = 192
    idx = [i for i in range(M) if i % 6 == 0][:32]      # one corner
    with Pool(8) as p:
        logs = p.map(one, [(cfg, M, i) for i in idx])
    solves = [l[0] for l in logs]
    n = np.array([len(s) for s in solves])
    print("attempts per instance: min %d mean %.0f max %d" % (n.min(), n.mean(), n.max()))
    print("solves per attempt: mean %.2f" % np.mean(np.concatenate(solves)))
    useful = sum(len(s) * (cpro + cepi) + s.sum() * clu for s in solves)
    # A: lock-step by attempt index, warp lives until the longest lane ends
    L = n.max()
    S = np.zeros((32, L))
    for k, s in enumerate(solves):
        S[k, :len(s)] = s
    costA = L * (cpro + cepi) + S.max(axis=0).sum() * clu
    print("A  lock-step, warp claims 32: efficiency %.3f" % (useful / (32 * costA)))
    # B: per-lane claiming: stream of attempts per lane, never idle (steady state)
    cat = np.concatenate(solves)
    rng = np.random.default_rng(0)
    trips = 20000
    draws = rng.choice(cat, size=(32, trips))
    costB = trips * (cpro + cepi) + draws.max(axis=0).sum() * clu
    usefulB = trips * 32 * (cpro + cepi) + draws.sum() * clu
    print("B  lock-step, per-lane claiming:  efficiency %.3f" % (usefulB / (32 * costB)))
    # C: flattened: every trip = one solve for every lane; prologue/epilogue run
    # (for the whole warp) whenever any lane needs them
    p_any = 1.0
    mean_s = cat.mean()
    usefulC = clu + (cpro + cepi) / mean_s
    costC = clu + p_any * (cpro + cepi)
    print("C  flattened, one solve per trip:  efficiency %.3f" % (usefulC / costC))
