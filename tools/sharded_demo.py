#!/usr/bin/env python
"""Multi-GPU check (run under torchrun): a cfg1 sweep sharded over the ranks with
eispice_b200.sharding, waveforms gathered over NCCL, compared on rank 0 with a
single-GPU run of the whole sweep."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
from eispice_b200 import engine, sharding  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
engine.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = workloads.make(sys.argv[1] if len(sys.argv) > 1 else "cfg1", int(sys.argv[2]) if len(sys.argv) > 2 else 4099)
res = sharding.tran_batch_sharded(w.cct, w.tstep, w.tstop, 0.0, w.params, w.probes)
if dist.get_rank() == 0:
    eng = engine.Engine(w.cct.netlist(), w.M)
    one = eng.tran_batch(w.tstep, w.tstop, 0.0, w.params, w.probes)
    same = np.array_equal(one.status, res.status)
    close = np.allclose(one.data, res.data, rtol=1e-9, atol=1e-12)
    print("world", dist.get_world_size(), "M", w.M, "shape", res.data.shape, "status equal", same,
          "waveforms equal", bool(np.array_equal(one.data, res.data)), "close", bool(close),
          "accepted", res.stats["accepted"], "vs", one.stats["accepted"])
    assert same and close
dist.barrier()
dist.destroy_process_group()
