"""Instance sharding across the GPUs of one box.

Instances of a sweep are independent, so the hot path needs no collective:
rank r of W owns the contiguous block ``shard_range(M, r, W)`` of the sweep,
runs it on its own GPU with its own engine, and only the finished probe
waveforms travel - ONE all_gather per run over NCCL (NVLink/NVSwitch) on GPUs,
gloo in the CPU tests.  One process per GPU, launched by torchrun.
"""
import numpy as np


def shard_range(num_instances, rank, world):
    """[start, stop) of rank's block: blocks of ceil(M / W), the last ones may
    be shorter or empty."""
    per = -(-int(num_instances) // int(world))
    start = min(num_instances, rank * per)
    return start, min(num_instances, start + per)


def shard_params(params, rank, world):
    """The rank's slice of every per-instance parameter array."""
    if not params:
        return {}
    M = max(len(np.atleast_1d(v)) for v in params.values())
    a, b = shard_range(M, rank, world)
    return {k: np.ascontiguousarray(np.broadcast_to(np.asarray(v), (M,))[a:b])
            for k, v in params.items()}


def gather_waveforms(local, num_instances, group=None, device=None):
    """All ranks pass their [m_r, ngrid, nprobes] block (m_r from shard_range);
    every rank gets the full [M, ngrid, nprobes] array back.  Blocks are padded
    to ceil(M / W) so a single all_gather moves them."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-int(num_instances) // world)
    a, b = shard_range(num_instances, rank, world)
    local = np.ascontiguousarray(local, dtype=np.float64)
    if local.shape[0] != b - a:
        raise ValueError("rank %d holds %d instances, expected %d" % (rank, local.shape[0], b - a))
    tail = local.shape[1:]
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) \
            if dist.get_backend(group) == "nccl" else torch.device("cpu")
    pad = torch.zeros((per,) + tail, dtype=torch.float64, device=device)
    if b > a:
        pad[:b - a].copy_(torch.from_numpy(local))
    out = torch.empty((world * per,) + tail, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:num_instances].cpu().numpy()


def tran_batch_sharded(circuit, tstep, tstop, tmax=0.0, params=None, probes=(), group=None,
                       **options):
    """``Circuit.tran_batch`` over all ranks of the process group: each rank
    simulates its shard on its current CUDA device; every rank returns the
    full-sweep waveforms [M, ngrid, nprobes] and the summed statistics."""
    import torch
    import torch.distributed as dist
    from . import engine, units
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    params = params or {}
    M = max(len(np.atleast_1d(v)) for v in params.values()) if params else 1
    a, b = shard_range(M, rank, world)
    mine = shard_params(params, rank, world)
    tstep, tstop, tmax = units.float(tstep), units.float(tstop), units.float(tmax)
    if b > a:
        eng = engine.Engine(circuit.netlist(), b - a)
        res = eng.tran_batch(tstep, tstop, tmax, mine, list(probes), **options)
        data, status, stats, ngrid = res.data, res.status, res.stats, eng.ngrid
        eng.close()
    else:
        data = status = stats = None
        ngrid = 0
    dev = torch.device("cuda", torch.cuda.current_device()) \
        if dist.get_backend(group) == "nccl" else torch.device("cpu")
    ng = torch.tensor([ngrid], device=dev)
    dist.all_reduce(ng, op=dist.ReduceOp.MAX, group=group)
    ngrid = int(ng.item())
    if data is None:
        data = np.zeros((0, ngrid, len(probes)))
        status = np.zeros(0, dtype=np.int32)
    full = gather_waveforms(data, M, group, dev)
    st = gather_waveforms(status.astype(np.float64).reshape(-1, 1, 1), M, group, dev)
    counters = torch.tensor([float(stats[k]) if stats else 0.0 for k in
                             ("accepted", "rejectedLTE", "rejectedNC", "factorizations", "failed")],
                            dtype=torch.float64, device=dev)
    dist.all_reduce(counters, group=group)
    total = dict(zip(("accepted", "rejectedLTE", "rejectedNC", "factorizations", "failed"),
                     [int(x) for x in counters.tolist()]))
    return engine.BatchResult(full, engine.output_grid(tstep, tstop, ngrid),
                              st.reshape(-1).astype(np.int32), total, list(probes))
