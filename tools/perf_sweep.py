#!/usr/bin/env python
"""Device-resident timing of one workload under several engine options.
usage: perf_sweep.py cfg M "opt=val,opt=val" ["..." ...]   (empty string = defaults)
Appends one JSON line per variant to gpurun_out/perf_sweep.jsonl.
Note: an explicit threads_per_block switches the automatic launch-shape policy
off, block_sync included - pass block_sync=1 with it to compare block sizes."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
from eispice_b200 import engine  # noqa: E402

name, M = sys.argv[1], int(sys.argv[2])
w = workloads.make(name, M)
n_tlines = sum(1 for d in w.cct.netlist() if d[0] == "T")
for spec in sys.argv[3:] or [""]:
    eng = engine.Engine(w.cct.netlist(), w.M)
    for kv in [x for x in spec.split(",") if x]:
        k, v = kv.split("=")
        eng.set_option(k, float(v))
    eng.set_params(w.params)
    t0 = time.time()
    try:
        eng.prepare(w.tstep, w.tstop, 0.0, w.probes)
        prep = time.time() - t0
        eng.run_device()
        eng.run_device()
        st = eng.stats()
        balg, k = workloads.algorithmic_bytes(st, n_tlines, len(w.probes))
        rate = st["accepted"] / (st["tranMs"] * 1e-3)
        line = {"cfg": name, "M": M, "opts": spec, "rate_M": round(rate / 1e6, 2), "ms": round(st["tranMs"], 1),
                "frac": round(balg * rate / 6547.8e9, 4), "failed": st["failed"], "regs": st["registersPerThread"],
                "local": st["localBytesPerThread"], "blocks": st["gridBlocks"], "tpb": st["threadsPerBlock"],
                "prepare_s": round(prep, 1), "k": round(k, 3)}
    except Exception as e:  # noqa: BLE001
        line = {"cfg": name, "M": M, "opts": spec, "error": str(e)}
    print(json.dumps(line), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "perf_sweep.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")
    eng.close()
