#!/usr/bin/env python
"""Static code-size attribution: SASS instructions of a cubin (compiled with
-lineinfo) per source function of the generated translation unit.
usage: sass_by_function.py file.cu file.cubin"""
import re
import subprocess
import sys
from collections import Counter

src_path, cubin = sys.argv[1], sys.argv[2]
src = open(src_path).read().splitlines()
func_at, cur = [], "?"
for ln in src:
    m = re.match(r'^(?:DEV|DEVFN|SIMCU_HD|extern "C" __global__)\s+.*?(\w+)\(', ln)
    if m:
        cur = m.group(1)
    func_at.append(cur)
out = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
cnt = Counter()
section = Counter()
line = None
sec = "?"
for ln in out.splitlines():
    m = re.search(r'//## File ".*?", line (\d+)', ln)
    if m:
        line = int(m.group(1))
        continue
    m = re.match(r"\s*\.section\s+(\S+)", ln)
    if m:
        sec = m.group(1)
    if sec.startswith(".text") and re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
        f = func_at[line - 1] if line and line <= len(func_at) else "?"
        cnt[f] += 1
        section[sec] += 1
tot = sum(cnt.values())
print("total SASS instructions: %d (%.0f KB)" % (tot, tot * 16 / 1024))
for s, n in section.most_common(12):
    print("  section %-60s %6d" % (s[:60], n))
for f, n in cnt.most_common(30):
    print("%-24s %6d  %5.1f%%" % (f, n, 100.0 * n / tot))
