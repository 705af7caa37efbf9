#!/bin/bash
# round-2 job 2: regression tests on the new kernel, claim / sync sweep, parity exploration
set -x
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02_j2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_j2_pytest.log
tail -5 gpurun_out/r02_j2_pytest.log
timeout 600 python tools/perf_sweep.py cfg5 131072 "claim_below=0" "claim_below=31" "claim_below=16" "claim_below=31,block_sync=1,threads_per_block=256" "claim_below=31,block_sync=1,threads_per_block=128" "claim_below=31,threads_per_block=256" > gpurun_out/r02_j2_sweep5.log 2>&1
timeout 300 python tools/perf_sweep.py cfg4 262144 "claim_below=0" "claim_below=31" "claim_below=31,block_sync=1,threads_per_block=256" > gpurun_out/r02_j2_sweep4.log 2>&1
timeout 120 python tools/perf_sweep.py cfg2 10000 "claim_below=0" "claim_below=31" > gpurun_out/r02_j2_sweep2.log 2>&1
cat gpurun_out/r02_j2_sweep*.log
for c in cfg1 cfg2 cfg3 cfg4 cfg5; do
  timeout 900 python tools/parity_explore.py $c 256 gpurun_out/r02_explore_$c.json > gpurun_out/r02_explore_$c.log 2>&1; echo "$c rc=$?"
done
tail -3 gpurun_out/r02_explore_*.log
