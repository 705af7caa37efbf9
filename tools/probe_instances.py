#!/usr/bin/env python
"""Run single sweep instances under option variants and print status / steps.
usage: probe_instances.py cfg M idx,idx,... "opt=val,..." ["..." ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
from eispice_b200 import engine  # noqa: E402

cfg, M = sys.argv[1], int(sys.argv[2])
idx = np.array([int(x) for x in sys.argv[3].split(",")])
w = workloads.make(cfg, M)
sub = {k: np.asarray(v)[idx] for k, v in w.params.items()}
for spec in sys.argv[4:] or [""]:
    eng = engine.Engine(w.cct.netlist(), len(idx))
    for kv in [x for x in spec.split(",") if x]:
        k, v = kv.split("=")
        eng.set_option(k, float(v))
    res = eng.tran_batch(w.tstep, w.tstop, 0.0, sub, w.probes)
    print(cfg, spec or "(default)", "status", res.status.tolist(), "accepted", eng.counters("accepted").tolist(),
          "rej_lte", eng.counters("rejected_lte").tolist(), "rej_nc", eng.counters("rejected_nc").tolist(), flush=True)
    eng.close()
