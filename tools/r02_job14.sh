#!/bin/bash
# round-2 job 14: final bench N=1 (default), reference arm, launch list, smoke
set -x
cd "$(dirname "$0")/.."
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
cat gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_bench_reference_n1.err; echo "ref rc=$?"
cat gpurun_out/r02_bench_reference_n1.json
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_plain_launch.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_bench_cfg5.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_launch.log 2>&1
tail -2 gpurun_out/r02_ncu_launch.log; wc -l gpurun_out/r02_launches_bench_cfg5.csv
