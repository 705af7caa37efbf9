#!/usr/bin/env python
"""Adds one ncu capture to profiles/summary.json (read by bench.py ->
roofline.traffic / roofline.issue_frac) and copies its raw metric page to
profiles/.  usage: ncu_summary.py <file.ncu-rep> <key e.g. cfg5@131072> <accepted steps per launch> <name>"""
import csv
import io
import json
import os
import subprocess
import sys

rep, key, accepted, name = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(h):
    v, u = m[h]
    x = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
    return x * scale


raw_csv = os.path.join(ROOT, "profiles", name + "_raw.csv")
with open(raw_csv, "w") as f:
    w = csv.writer(f)
    for h in hdr:
        if h in m and not h.startswith("ID"):
            w.writerow([h, m[h][1], m[h][0]])
if "dram__bytes_read.sum" in m:
    dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    how = "dram__bytes_read.sum + dram__bytes_write.sum"
else:      # captures without the memory section: average rate x duration
    rate = float(m["dram__bytes.sum.per_second"][0]) * {"Gbyte/s": 1e9, "Tbyte/s": 1e12, "Mbyte/s": 1e6}[m["dram__bytes.sum.per_second"][1]]
    dur = float(m["gpu__time_duration.sum"][0]) * {"s": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9}[m["gpu__time_duration.sum"][1]]
    dram = rate * dur
    how = "dram__bytes.sum.per_second x gpu__time_duration.sum"
inst = num("smsp__inst_executed.sum")
path = os.path.join(ROOT, "profiles", "summary.json")
summ = json.load(open(path)) if os.path.exists(path) else {}
summ[key] = {
    "dram_bytes_per_launch": dram,
    "warp_instructions_per_launch": inst,
    "accepted_instance_timesteps_per_launch": accepted,
    "warp_instructions_per_instance_timestep": inst / accepted,
    "source": "profiles/%s_raw.csv (ncu, %s and smsp__inst_executed.sum, one launch of "
              "simcuJitKernel)" % (name, how),
}
json.dump(summ, open(path, "w"), indent=1)
print(json.dumps(summ[key], indent=1))
