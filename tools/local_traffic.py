#!/usr/bin/env python
"""Static local-memory traffic of the specialised kernel, without a GPU: generates
and compiles the kernel of a workload (NVRTC source, nvcc for the cubin with line
information), then counts LDL / STL bytes per thread in the SASS, attributed to
the statement of simcuJitKernel they were inlined into (the phase) - or, with a
third argument, to the innermost source line inside ONE phase.

usage: local_traffic.py cfg5 896 ["opt=val,opt=val" ["text of a kernel statement"]]
e.g.   local_traffic.py cfg5 896 "solve_blocks=0,shared_break_clock=0"
       local_traffic.py cfg5 896 "" "brk = genStep(c);"

Every path is counted once (both sides of a branch), so the numbers are an upper
bound per call and a ranking, not a measurement: profiles/r02b_ncu_*.md sets them
beside ncu's executed local loads / stores."""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
from eispice_b200 import engine  # noqa: E402


def build(name, tpb, spec, work):
    w = workloads.make(name, 64)
    eng = engine.Engine(w.cct.netlist(), w.M)
    eng.set_option("threads_per_block", tpb)
    eng.set_option("block_sync", 1)
    for kv in [x for x in spec.split(",") if x]:
        k, v = kv.split("=")
        eng.set_option(k, float(v))
    eng.set_params(w.params)
    src, _ = eng.compile_only(w.probes)
    eng.close()
    cu = os.path.join(work, "kernel.cu")
    cubin = os.path.join(work, "kernel.cubin")
    open(cu, "w").write(src)
    fmad = "--fmad=true" if "fmad=1" in spec else "--fmad=false"
    out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", fmad,
                          "-lineinfo", "-cubin", "-o", cubin, "-Xptxas", "-v", cu],
                         capture_output=True, text=True)
    if out.returncode:
        raise SystemExit(out.stderr)
    for ln in out.stderr.splitlines():
        if "registers" in ln or "stack frame" in ln:
            print(ln.strip())
            if "registers" in ln:
                break
    return cu, cubin


def walk(cu, cubin):
    """Yields (outermost kernel line, innermost line of the generated file, opcode text)."""
    out = subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True,
                         text=True).stdout
    sec, chain, pending = "?", [], []
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+(\S+)", ln)
        if m:
            sec = m.group(1)
            continue
        m = re.search(r'//## File "(.*?)", line (\d+)( inlined at "(.*?)", line (\d+))?', ln)
        if m:
            pending.append((m.group(1), int(m.group(2)), m.group(4)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln) and sec.startswith(".text.simcuJitKernel"):
            if pending:
                chain, pending = pending, []
            outer = [c for c in chain if c[2] is None]
            kline = outer[-1][1] if outer and outer[-1][0] == cu else 0
            inner = chain[0][1] if chain and chain[0][0] == cu else 0
            yield kline, inner, ln


def width(text):
    m = re.search(r"\b(LDL|STL)(\.\w+)*", text)
    if not m:
        return None, 0
    w = 8 if ".64" in m.group(0) else (16 if ".128" in m.group(0) else 4)
    return m.group(1), w


def main():
    name, tpb = sys.argv[1], int(sys.argv[2])
    spec = sys.argv[3] if len(sys.argv) > 3 else ""
    only = sys.argv[4] if len(sys.argv) > 4 else None
    with tempfile.TemporaryDirectory() as work:
        cu, cubin = build(name, tpb, spec, work)
        src = open(cu).read().splitlines()
        want = None
        if only:
            want = next(i for i, l in enumerate(src) if only in l) + 1
        ins, ld, st = Counter(), Counter(), Counter()
        for kline, inner, text in walk(cu, cubin):
            if want is not None and kline != want:
                continue
            key = inner if want is not None else kline
            ins[key] += 1
            kind, w = width(text)
            if kind == "LDL":
                ld[key] += w
            elif kind == "STL":
                st[key] += w
        print("%d instructions, %d B of local loads, %d B of local stores (per thread, all paths)"
              % (sum(ins.values()), sum(ld.values()), sum(st.values())))
        for key, _ in sorted(ins.items(), key=lambda kv: -(ld[kv[0]] + st[kv[0]]))[:30]:
            text = src[key - 1].strip()[:76] if key > 0 else "?"
            print("%5d  instr %5d  LDL %5d  STL %5d | %s" % (key, ins[key], ld[key], st[key], text))


if __name__ == "__main__":
    main()
