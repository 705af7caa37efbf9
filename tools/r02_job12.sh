#!/bin/bash
# round-2 job 12 (2 GPUs): the whole GPU test suite on GPU 0 (sharded tests use both), then bench N=2 with the pipelined gather
set -x
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02_j12_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_j12_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29756 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench2 rc=$?"
cat gpurun_out/r02_bench_n2.json; tail -3 gpurun_out/r02_bench_n2.err
