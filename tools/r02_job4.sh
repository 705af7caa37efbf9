#!/bin/bash
# round-2 job 4: parity exploration (fixed), occupancy variants, guard on/off
set -x
cd "$(dirname "$0")/.."
for c in cfg1 cfg2 cfg3 cfg4 cfg5; do
  timeout 900 python tools/parity_explore.py $c 256 gpurun_out/r02_explore_$c.json > gpurun_out/r02_explore_$c.log 2>&1; echo "$c rc=$?"
done
timeout 900 python tools/perf_sweep.py cfg5 131072 "block_sync=1,threads_per_block=512" "block_sync=1,threads_per_block=512,growth_limit=0" "block_sync=1,threads_per_block=768" "block_sync=1,threads_per_block=1024" "block_sync=1,threads_per_block=512,claim_below=0" > gpurun_out/r02_j4_sweep5.log 2>&1
timeout 300 python tools/perf_sweep.py cfg4 262144 "block_sync=1,threads_per_block=256" "block_sync=1,threads_per_block=256,growth_limit=0" "block_sync=1,threads_per_block=384" > gpurun_out/r02_j4_sweep4.log 2>&1
timeout 300 python tools/perf_sweep.py cfg3 100000 "" "block_sync=1,threads_per_block=256" > gpurun_out/r02_j4_sweep3.log 2>&1
timeout 300 python tools/perf_sweep.py cfg2 10000 "" "block_sync=1,threads_per_block=256" > gpurun_out/r02_j4_sweep2.log 2>&1
cat gpurun_out/r02_j4_sweep*.log
