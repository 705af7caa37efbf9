"""The multi-GPU plumbing on CPU: world_size-2 gloo process group, no CUDA.
Instances are dealt out round-robin with no data-path collective; one gather to
the root moves the waveforms (NCCL between device buffers on GPUs - see
csrc/simcu_comm.cpp and tools/sharded_demo.py; here the same index logic over
gloo host blocks)."""
import os
import socket

import numpy as np
import pytest

from eispice_b200 import sharding


def test_round_robin_shards_partition_the_sweep():
    for M in (1, 7, 64, 1000, 262144):
        for W in (1, 2, 3, 8):
            idx = [sharding.shard_indices(M, r, W) for r in range(W)]
            assert sorted(np.concatenate(idx).tolist()) == list(range(M))
            per = sharding.shard_size(M, W)
            assert all(len(i) <= per for i in idx) and per * W >= M
            for r, i in enumerate(idx):
                assert (i % W == r).all()


def test_shard_params_slices_pads_and_broadcasts():
    p = {"R1.R": np.arange(10.0), "C1.C": np.array([3.0])}
    got = [sharding.shard_params(p, r, 4) for r in range(4)]
    assert [len(g["R1.R"]) for g in got] == [3, 3, 3, 3]        # padded to ceil(10 / 4)
    assert got[0]["R1.R"].tolist() == [0.0, 4.0, 8.0]
    assert got[2]["R1.R"].tolist() == [2.0, 6.0, 6.0]           # pad repeats the last own instance
    assert all((g["C1.C"] == 3.0).all() for g in got)
    # interleaving the padded blocks restores the sweep
    full = sharding.interleave([g["R1.R"] for g in got], 10)
    assert np.array_equal(full, np.arange(10.0))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, M, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = sharding.Gatherer()
        mine = sharding.shard_params({"x": np.arange(M, dtype=np.float64)}, rank, world, M)["x"]
        # each instance's "waveform" encodes its global index
        local = mine[:, None, None] * np.ones((1, 5, 2))
        local[:, :, 1] += 0.5
        stat = (mine.astype(np.int32) % 3)
        full, st = g.gather_host(local, stat, M, root=0)
        if rank == 0:
            ok = (full.shape == (M, 5, 2) and np.array_equal(full[:, 0, 0], np.arange(M))
                  and np.array_equal(full[:, 3, 1], np.arange(M) + 0.5)
                  and np.array_equal(st, np.arange(M) % 3))
        else:
            ok = full is None and st is None
        piv = sharding.broadcast_array(np.arange(6, dtype=np.int32) * (rank == 0), 0)
        ok = ok and piv.tolist() == [0, 1, 2, 3, 4, 5]
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("M", [9, 2, 1])
def test_gather_to_root_over_gloo_world2(M):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, M, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
