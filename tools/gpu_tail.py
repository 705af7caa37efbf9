#!/usr/bin/env python
"""Development probe: distribution of per-instance work in a batch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads
from eispice_b200 import engine
name, M = sys.argv[1], int(sys.argv[2])
w = workloads.make(name, M)
eng = engine.Engine(w.cct.netlist(), w.M)
eng.set_params(w.params)
eng.prepare(w.tstep, w.tstop, 0.0, w.probes)
eng.upload(); eng.run_device()
st = eng.stats()
f = eng.counters("factorizations").astype(np.int64)
a = eng.counters("accepted").astype(np.int64)
s = eng.counters("status")
print(name, M, "ms", st["tranMs"], "fact mean", f.mean(), "max", f.max(), "p50", np.percentile(f, 50),
      "p99", np.percentile(f, 99), "p99.9", np.percentile(f, 99.9), "max/mean", f.max() / f.mean())
print(" accepted mean", a.mean(), "max", a.max(), "failed", (s != 0).sum())
order = np.argsort(-f)[:12]
print(" top instances", [(int(i), int(f[i]), int(a[i]), int(s[i]), int(i) % 6) for i in order])
np.save(os.path.join(ROOT, "gpurun_out", "fact_%s.npy" % name), f)
