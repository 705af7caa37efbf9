#!/usr/bin/env python
"""Development probe: raw GPU record + pivots of chosen instances of a workload
batch -> gpurun_out/dump_<cfg>.npz.  Usage: gpu_dump_instance.py cfg4 36 35 [more]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
from eispice_b200 import engine  # noqa: E402

name, M = sys.argv[1], int(sys.argv[2])
inst = [int(x) for x in sys.argv[3:]]
w = workloads.make(name, M)
eng = engine.Engine(w.cct.netlist(), w.M)
res = eng.tran_batch(w.tstep, w.tstop, 0.0, w.params, w.probes, raw_max_points=40000,
                     raw_count=w.M)
out = {"status": res.status, "names": np.array(eng.variables())}
for ph in (0, 1):
    pc, pr = eng.pivots(ph)
    out["pc%d" % ph], out["pr%d" % ph] = pc, pr
for i in inst:
    out["raw%d" % i] = eng.fetch_raw(i)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "dump_%s.npz" % name), **out)
print(res.status, res.stats)
