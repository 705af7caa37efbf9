"""Two ranks on two GPUs (skipped on a single-GPU box): the sweep sharded
round-robin, gathered to rank 0 over NCCL between device buffers
(csrc/simcu_comm.cpp), bit-equal to the single-GPU run - tools/sharded_demo.py
under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("cfg,M", [("cfg4", 4099), ("cfg1", 513)])
def test_sharded_sweep_equals_single_gpu(cfg, M):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tools", "sharded_demo.py"), cfg, str(M)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert '"waveforms_bit_equal": true' in out.stdout
