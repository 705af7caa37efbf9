#!/usr/bin/env python
"""Offline (no GPU): generate + NVRTC-compile the specialised kernel of each
workload and report ptxas registers / stack / spills via nvcc on the source."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
from eispice_b200 import engine  # noqa: E402

names = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
for name in names:
    w = workloads.make(name, 64)
    e = engine.Engine(w.cct.netlist(), w.M)
    e.set_params(w.params)
    t = time.time()
    src, n = e.compile_only(w.probes)
    dt = time.time() - t
    path = "/tmp/gen_%s.cu" % name
    open(path, "w").write(src)
    out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17",
                          "-cubin", "-o", path.replace(".cu", ".cubin"), "-Xptxas", "-v", path],
                         capture_output=True, text=True).stderr
    lines = [ln.strip() for ln in out.splitlines() if "registers" in ln or "spill" in ln]
    print(name, "nvrtc %.1fs" % dt, "cubin %d B" % n, "|", " ; ".join(lines[:2]))
