#!/usr/bin/env python
"""One batched transient of a workload, for ncu: prepare, one warm-up run, one
profiled run.  Usage: profile_run.py <cfg> <instances> ["opt=val,opt=val"]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
from eispice_b200 import engine  # noqa: E402

name, M = sys.argv[1], int(sys.argv[2])
w = workloads.make(name, M)
eng = engine.Engine(w.cct.netlist(), w.M)
for kv in [x for x in (sys.argv[3] if len(sys.argv) > 3 else "").split(",") if x]:
    k, v = kv.split("=")
    eng.set_option(k, float(v))
eng.set_params(w.params)
eng.prepare(w.tstep, w.tstop, 0.0, w.probes)
eng.upload()
for _ in range(2):
    eng.run_device()
st = eng.stats()
print({k: st[k] for k in ("accepted", "factorizations", "failed", "deviceMs", "tranMs",
                          "threadsPerBlock", "numFactor")})
