#!/bin/bash
# round-2 job 5: regression, auto launch shape + T-line prefetch pass, replay with convergence-test outcomes
set -x
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02_j5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_j5_pytest.log
tail -15 gpurun_out/r02_j5_pytest.log
timeout 900 python tools/perf_sweep.py cfg5 131072 "" "tline_prefetch=0" "threads_per_block=896,block_sync=1" "threads_per_block=512,block_sync=1" "threads_per_block=1024,block_sync=0" > gpurun_out/r02_j5_sweep5.log 2>&1
timeout 300 python tools/perf_sweep.py cfg4 262144 "" > gpurun_out/r02_j5_sweep4.log 2>&1
cat gpurun_out/r02_j5_sweep*.log
for c in cfg2 cfg4 cfg5; do
  timeout 900 python tools/parity_explore.py $c 256 gpurun_out/r02_explore2_$c.json > gpurun_out/r02_explore2_$c.log 2>&1; echo "$c rc=$?"
done
timeout 1500 python -m pytest tests/test_gpu_replay.py -x -q -s > gpurun_out/r02_j5_replay.log 2>&1; echo "replay rc=$?"
tail -30 gpurun_out/r02_j5_replay.log
