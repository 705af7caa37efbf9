#!/bin/bash
# round-2 job 10 (2 GPUs): sharded sweep == single GPU, bench --gpus 2 with the NCCL gather in the e2e timer
set -x
cd "$(dirname "$0")/.."
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "bit_identical or ten_decades" > gpurun_out/r02_j10_bit.log 2>&1; tail -3 gpurun_out/r02_j10_bit.log
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -s > gpurun_out/r02_j10_sharded.log 2>&1; echo "sharded rc=$?"; tail -15 gpurun_out/r02_j10_sharded.log
cat gpurun_out/sharded_demo.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29755 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench2 rc=$?"
cat gpurun_out/r02_bench_n2.json; tail -5 gpurun_out/r02_bench_n2.err
free -g | head -2
