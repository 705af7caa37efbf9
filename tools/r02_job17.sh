#!/bin/bash
set -x
cd "$(dirname "$0")/.."
timeout 300 python tools/check_smem_tables.py > gpurun_out/r02_j17_check.log 2>&1; echo "check rc=$?"; tail -3 gpurun_out/r02_j17_check.log
timeout 400 python tools/perf_sweep.py cfg4 262144 "" "smem_tables_kb=100" "smem_tables_kb=200" > gpurun_out/r02_j17_sweep4.log 2>&1
timeout 600 python tools/perf_sweep.py cfg5 131072 "" "smem_tables_kb=100" > gpurun_out/r02_j17_sweep5.log 2>&1
cat gpurun_out/r02_j17_sweep*.log
