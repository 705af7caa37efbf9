#!/usr/bin/env python
"""Tables staged in shared memory must not change a single bit of the result."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
from eispice_b200 import engine  # noqa: E402

for cfg, M in (("cfg4", 4096), ("cfg5", 1024)):
    w = workloads.make(cfg, M)
    out = []
    for kb in (0, 100, 200):
        eng = engine.Engine(w.cct.netlist(), w.M)
        res = eng.tran_batch(w.tstep, w.tstop, 0.0, w.params, w.probes, threads_per_block=256,
                             block_sync=1, smem_tables_kb=kb)
        out.append((res.data.copy(), res.status.copy()))
        eng.close()
    same = all(np.array_equal(out[0][0], o[0], equal_nan=True) and np.array_equal(out[0][1], o[1]) for o in out[1:])
    print(cfg, M, "bit-identical with staged tables:", same, "failed", int((out[0][1] != 0).sum()))
    assert same
