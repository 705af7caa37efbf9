#!/bin/bash
# round-2 baseline of the round-1 kernel at BASELINE sizes: timing, then ncu (reduced sections)
set -x
SEC="--section SpeedOfLight --section Occupancy --section WarpStateStats --section SourceCounters --section MemoryWorkloadAnalysis --section LaunchStats --section InstructionStats --section SchedulerStats"
python tools/profile_run.py cfg4 262144 > gpurun_out/r02_base_cfg4.log 2>&1 &&
timeout 900 ncu $SEC --clock-control none --import-source on -k simcuJitKernel -s 1 -c 1 -f -o gpurun_out/r02_base_cfg4_262144 python tools/profile_run.py cfg4 262144 > gpurun_out/r02_base_cfg4_ncu.log 2>&1
python tools/profile_run.py cfg5 131072 > gpurun_out/r02_base_cfg5.log 2>&1 &&
timeout 1200 ncu $SEC --clock-control none --import-source on -k simcuJitKernel -s 1 -c 1 -f -o gpurun_out/r02_base_cfg5_131072 python tools/profile_run.py cfg5 131072 > gpurun_out/r02_base_cfg5_ncu.log 2>&1
tail -3 gpurun_out/r02_base_cfg4.log gpurun_out/r02_base_cfg5.log
ls -la gpurun_out/*.ncu-rep
