"""The multi-GPU plumbing on CPU: world_size-2 gloo process group, no CUDA.
Instances shard with no data-path collective; one all_gather moves waveforms."""
import os
import socket

import numpy as np
import pytest

from eispice_b200 import sharding


def test_shard_ranges_partition_the_sweep():
    for M in (1, 7, 64, 1000, 262144):
        for W in (1, 2, 3, 8):
            spans = [sharding.shard_range(M, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == M
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
            assert sum(b - a for a, b in spans) == M


def test_shard_params_slices_and_broadcasts():
    p = {"R1.R": np.arange(10.0), "C1.C": np.array([3.0])}
    got = [sharding.shard_params(p, r, 4) for r in range(4)]
    assert np.array_equal(np.concatenate([g["R1.R"] for g in got]), np.arange(10.0))
    assert all((g["C1.C"] == 3.0).all() for g in got)
    assert [len(g["R1.R"]) for g in got] == [3, 3, 3, 1]


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
        a, b = sharding.shard_range(M, rank, world)
        # each instance's "waveform" encodes its global index
        local = np.arange(a, b, dtype=np.float64)[:, None, None] * np.ones((1, 5, 2))
        local[:, :, 1] += 0.5
        full = sharding.gather_waveforms(local, M)
        ok = (full.shape == (M, 5, 2) and np.array_equal(full[:, 0, 0], np.arange(M))
              and np.array_equal(full[:, 3, 1], np.arange(M) + 0.5))
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("M", [9, 2, 1])
def test_gather_over_gloo_world2(M):
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
