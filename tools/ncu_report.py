#!/usr/bin/env python3
"""Summarise an ncu capture of the specialised kernel: headline metrics, stall
reasons, and samples / instructions by source function.
usage: ncu_report.py <file.ncu-rep> [out.md]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]


def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout


out = []
rows = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
out.append("| metric | value | unit |\n|---|---|---|")
stalls = []
for h, u, v in zip(hdr, units, vals):
    if h in keys:
        out.append("| %s | %s | %s |" % (h, v, u))
    if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        try:
            stalls.append((float(v), h.replace("smsp__average_warps_issue_stalled_", "")
                           .replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
stalls.sort(reverse=True)
out.append("\nWarp stall reasons (warps stalled per issue-active cycle):\n")
out.append("| reason | ratio |\n|---|---|")
for v, h in stalls[:8]:
    out.append("| %s | %.2f |" % (h, v))

rows = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source", "cuda,sass"]))))
if len(rows) > 3 and rows[0][0] == "File Path":
    path = rows[0][1]
    hdr = rows[2]
    S, I, TI = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")

    def num(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    lines = [r for r in rows[3:] if r and r[0] not in ("",)]
    tot = sum(num(r[S]) for r in lines)
    toti = sum(num(r[I]) for r in lines)
    import os
    cands = [path, path.replace("/root/repo/", ""),
             os.path.join("gpurun_out", "src", os.path.basename(path)),
             os.path.join("profiles", "src", os.path.basename(path))]
    src = None
    for cnd in cands:
        if os.path.exists(cnd):
            src = open(cnd).read().splitlines()
            break
    if src is None:
        raise SystemExit("source %s not found (run with SIMCU_JIT_DUMP=gpurun_out/src)" % path)
    func_at, cur = [], "?"
    for ln in src:
        m = re.match(r'^(?:DEV|DEVFN|SIMCU_HD|extern "C" __global__)\s+.*?(\w+)\(', ln)
        if m:
            cur = m.group(1)
        func_at.append(cur)
    sc = {h: i for i, h in enumerate(hdr) if h in ("stall_no_inst", "stall_long_sb", "stall_wait",
                                                   "stall_short_sb", "stall_branch_resolving")}
    agg = {}
    for r in lines:
        f = func_at[int(r[0]) - 1]
        a = agg.setdefault(f, [0, 0, 0, {}])
        a[0] += num(r[S]); a[1] += num(r[I]); a[2] += num(r[TI])
        for h, i in sc.items():
            a[3][h] = a[3].get(h, 0) + num(r[i])
    out.append("\nBy source function (stall samples, executed warp instructions, active threads, top stalls):\n")
    out.append("| function | samples | instructions | avg threads | top stalls |\n|---|---|---|---|---|")
    for f, (s_, i_, t_, stl) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
        top = sorted(stl.items(), key=lambda kv: -kv[1])[:2]
        out.append("| %s | %.1f%% | %.1f%% | %.1f | %s |" % (
            f, 100 * s_ / tot, 100 * i_ / toti, t_ / max(i_, 1),
            ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100 * v / max(s_, 1)) for k, v in top)))
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
