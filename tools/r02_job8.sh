#!/bin/bash
# round-2 job 8: bench N=1 (default = cfg5@131072), smoke, ncu at BASELINE sizes (application replay)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/src
timeout 300 python tools/probe_instances.py cfg4 262144 52883,8781,100 "" "fmad=0" "pivot_threshold=0.5" "fmad=0,pivot_threshold=0.5" "max_attempts=200000" > gpurun_out/r02_probe4.log 2>&1; cat gpurun_out/r02_probe4.log
timeout 300 python tools/probe_instances.py cfg5 131072 61943,21953,100 "" "fmad=0" "fmad=0,pivot_threshold=0.5" > gpurun_out/r02_probe5.log 2>&1; cat gpurun_out/r02_probe5.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
cat gpurun_out/r02_bench_n1.json; tail -5 gpurun_out/r02_bench_n1.err
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
export SIMCU_JIT_DUMP=gpurun_out/src
SEC="--section SpeedOfLight --section Occupancy --section WarpStateStats --section SourceCounters --section MemoryWorkloadAnalysis --section LaunchStats --section InstructionStats --section SchedulerStats"
python tools/profile_run.py cfg5 131072 > gpurun_out/r02_prof_cfg5.log 2>&1 &&
timeout 1500 ncu $SEC --replay-mode application --clock-control none --import-source on -k simcuJitKernel -s 1 -c 1 -f -o gpurun_out/r02_cfg5_131072 python tools/profile_run.py cfg5 131072 > gpurun_out/r02_ncu_cfg5.log 2>&1
tail -3 gpurun_out/r02_ncu_cfg5.log
python tools/profile_run.py cfg4 262144 > gpurun_out/r02_prof_cfg4.log 2>&1 &&
timeout 900 ncu $SEC --replay-mode application --clock-control none --import-source on -k simcuJitKernel -s 1 -c 1 -f -o gpurun_out/r02_cfg4_262144 python tools/profile_run.py cfg4 262144 > gpurun_out/r02_ncu_cfg4.log 2>&1
tail -3 gpurun_out/r02_ncu_cfg4.log
ls -la gpurun_out/*.ncu-rep
