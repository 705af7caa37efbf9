#!/bin/bash
# round-2 job 7: replay + full-size tests with rescue, perf variants with trimmed break-point state
set -x
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests/test_gpu_replay.py -q -s > gpurun_out/r02_j7_replay.log 2>&1; echo "replay rc=$?"
grep -E "replayed|under SuperLU|passed|failed|Error" gpurun_out/r02_j7_replay.log | head -30
timeout 1800 python -m pytest tests/test_gpu_fullsize.py -q -s -k "cfg4 or cfg5" > gpurun_out/r02_j7_fullsize.log 2>&1; echo "fullsize rc=$?"
grep -E "within 1x|flagged|passed|failed|Error" gpurun_out/r02_j7_fullsize.log | head -30
timeout 900 python tools/perf_sweep.py cfg5 131072 "" "threads_per_block=512,min_blocks_per_sm=2,block_sync=1" "threads_per_block=896,block_sync=1" > gpurun_out/r02_j7_sweep5.log 2>&1
timeout 300 python tools/perf_sweep.py cfg4 262144 "" > gpurun_out/r02_j7_sweep4.log 2>&1
cat gpurun_out/r02_j7_sweep*.log
