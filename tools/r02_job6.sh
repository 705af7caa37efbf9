#!/bin/bash
# round-2 job 6: replay with the oracle's arithmetic (fmad off), full-size adaptive tests
set -x
cd "$(dirname "$0")/.."
for c in cfg2 cfg4 cfg5; do
  timeout 900 python tools/parity_explore.py $c 256 gpurun_out/r02_explore3_$c.json > gpurun_out/r02_explore3_$c.log 2>&1; echo "$c rc=$?"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ten_decades or one_call" > gpurun_out/r02_j6_pytest.log 2>&1; tail -5 gpurun_out/r02_j6_pytest.log
timeout 1800 python -m pytest tests/test_gpu_fullsize.py -q -s > gpurun_out/r02_j6_fullsize.log 2>&1; echo "fullsize rc=$?"
grep -E "within 1x|flagged|passed|failed|Error|assert" gpurun_out/r02_j6_fullsize.log | head -40
