#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU): a sweep dealt out
round-robin over the ranks with eispice_b200.sharding, waveforms gathered to
rank 0 over NCCL between device buffers (simcuGatherOutput), compared on rank 0
with a single-GPU run of the whole sweep under the same pivot sequences:
bit-equal.  usage: sharded_demo.py [cfg] [instances]"""
import json
import os
import sys
import time

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
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4099
w = workloads.make(cfg, M)
g = sharding.Gatherer()
t0 = time.time()
res = sharding.tran_batch_sharded(w.cct, w.tstep, w.tstop, 0.0, w.params, w.probes, gatherer=g)
dt = time.time() - t0
if dist.get_rank() == 0:
    # the whole sweep on ONE GPU with the pivot sequences the sharded run used:
    # rank 0 derived them on its shard's pilots and broadcast them
    W = dist.get_world_size()
    mine = sharding.shard_params(w.params, 0, W, w.M)
    e0 = engine.Engine(w.cct.netlist(), sharding.shard_size(w.M, W))
    e0.set_params(mine)
    e0.prepare(w.tstep, w.tstop, 0.0, w.probes)
    piv = e0.pivots(0) + e0.pivots(1)
    e0.close()
    eng = engine.Engine(w.cct.netlist(), w.M)
    eng.set_pivots(0, piv[0], piv[1])
    eng.set_pivots(1, piv[2], piv[3])
    one = eng.tran_batch(w.tstep, w.tstop, 0.0, w.params, w.probes)
    ok = one.status == 0
    line = {"world": W, "cfg": cfg, "M": w.M, "shape": list(res.data.shape),
            "status_equal": bool(np.array_equal(one.status, res.status)),
            "waveforms_bit_equal": bool(np.array_equal(one.data[ok], res.data[ok])),
            "accepted": [int(res.stats["accepted"]), int(one.stats["accepted"])],
            "gather": res.stats["gather"], "wall_s": round(dt, 3)}
    print(json.dumps(line))
    out = os.path.join(ROOT, "gpurun_out", "sharded_demo.jsonl")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "a") as f:
        f.write(json.dumps(line) + "\n")
    assert line["status_equal"] and line["waveforms_bit_equal"]
    assert res.stats["accepted"] == one.stats["accepted"]
g.close()
dist.barrier()
dist.destroy_process_group()
