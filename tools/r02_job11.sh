#!/bin/bash
set -x
cd "$(dirname "$0")/.."
timeout 900 python tools/perf_sweep.py cfg5 131072 "" "sync_group=7" "sync_group=14" "sync_group=4" > gpurun_out/r02_j11_sweep5.log 2>&1
timeout 400 python tools/perf_sweep.py cfg4 262144 "" "sync_group=4" "sync_group=2" > gpurun_out/r02_j11_sweep4.log 2>&1
cat gpurun_out/r02_j11_sweep*.log
