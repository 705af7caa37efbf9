#!/bin/bash
# round-2 job 9: fmad on/off at full size, full-size tests with the fmad rescue, cfg5@131072 ncu (kernel replay, short ring)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/src
timeout 600 python tools/perf_sweep.py cfg5 131072 "" "fmad=0" > gpurun_out/r02_j9_sweep5.log 2>&1
timeout 300 python tools/perf_sweep.py cfg4 262144 "" "fmad=0" > gpurun_out/r02_j9_sweep4.log 2>&1
timeout 120 python tools/perf_sweep.py cfg2 10000 "" "fmad=0" > gpurun_out/r02_j9_sweep2.log 2>&1
cat gpurun_out/r02_j9_sweep*.log
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -q -s -k "cfg4 or cfg5" > gpurun_out/r02_j9_fullsize.log 2>&1; echo "fullsize rc=$?"
grep -E "within 1x|flagged|passed|failed|Error" gpurun_out/r02_j9_fullsize.log | head -20
export SIMCU_JIT_DUMP=gpurun_out/src
python tools/profile_run.py cfg5 131072 "hist_records=1024" > gpurun_out/r02_j9_prof5.log 2>&1 &&
timeout 700 ncu --section SpeedOfLight --section WarpStateStats --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none -k simcuJitKernel -s 1 -c 1 -f -o gpurun_out/r02_cfg5_131072 python tools/profile_run.py cfg5 131072 "hist_records=1024" > gpurun_out/r02_j9_ncu5.log 2>&1
cat gpurun_out/r02_j9_prof5.log; tail -2 gpurun_out/r02_j9_ncu5.log
