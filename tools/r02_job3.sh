#!/bin/bash
# round-2 job 3: parity exploration, sweep with the angle filter, source-level ncu of cfg5 at one full wave
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/src
for c in cfg1 cfg2 cfg3 cfg4 cfg5; do
  timeout 900 python tools/parity_explore.py $c 256 gpurun_out/r02_explore_$c.json > gpurun_out/r02_explore_$c.log 2>&1; echo "$c rc=$?"
done
timeout 900 python tools/perf_sweep.py cfg5 131072 "" "block_sync=1,threads_per_block=256" "block_sync=1,threads_per_block=256,min_blocks_per_sm=2" "block_sync=1,threads_per_block=512" "block_sync=1,threads_per_block=256,claim_below=0" > gpurun_out/r02_j3_sweep5.log 2>&1
timeout 300 python tools/perf_sweep.py cfg4 262144 "" "block_sync=1,threads_per_block=256" "block_sync=1,threads_per_block=256,min_blocks_per_sm=2" > gpurun_out/r02_j3_sweep4.log 2>&1
cat gpurun_out/r02_j3_sweep*.log
export SIMCU_JIT_DUMP=gpurun_out/src
SEC="--section SpeedOfLight --section WarpStateStats --section SourceCounters --section SchedulerStats --section LaunchStats --section InstructionStats"
python tools/profile_run.py cfg5 37888 "block_sync=1,threads_per_block=256" > gpurun_out/r02_j3_prof5.log 2>&1 &&
timeout 600 ncu $SEC --clock-control none --import-source on -k simcuJitKernel -s 1 -c 1 -f -o gpurun_out/r02_cfg5_37888_sync python tools/profile_run.py cfg5 37888 "block_sync=1,threads_per_block=256" > gpurun_out/r02_j3_ncu5.log 2>&1
cat gpurun_out/r02_j3_prof5.log
ls -la gpurun_out/*.ncu-rep gpurun_out/src
