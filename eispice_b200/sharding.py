"""Instance sharding across the GPUs of one box.

Instances of a sweep are independent, so the hot path needs no collective.
One process per GPU (torchrun).  The sweep is dealt out ROUND-ROBIN: global
instance g runs on rank ``g % W`` as its local instance ``g // W``, so every
rank holds the same mix of IBIS corners / parameter ranges and the ranks finish
together.  Shards are padded to ``ceil(M / W)`` instances (the pad repeats the
rank's last instance and is dropped again after the gather), so every rank runs
the same shape.

Only the finished probe waveforms travel - ONE gather to the root rank per run:

* NCCL (GPUs): ``simcuGatherOutput`` of the C-ABI (csrc/simcu_comm.cpp) -
  grouped ncclSend / ncclRecv between DEVICE buffers over NVLink / NVSwitch,
  then one strided device -> host copy per rank on the root, which also
  de-interleaves.  torch.distributed only bootstraps it (broadcast of the
  128-byte NCCL id, of the pivot sequences and of a few counters).
* gloo (the CPU tests of this host logic): ``dist.gather`` of host blocks.
"""
import ctypes as C

import numpy as np


def shard_size(num_instances, world):
    """Instances per rank after padding: ceil(M / W)."""
    return -(-int(num_instances) // int(world))


def shard_indices(num_instances, rank, world):
    """Global indices of the instances rank owns (round-robin), unpadded."""
    return np.arange(int(rank), int(num_instances), int(world))


def shard_params(params, rank, world, num_instances=None):
    """The rank's slice of every per-instance parameter array, padded to
    shard_size by repeating the rank's last instance (instance 0 of the sweep
    for a rank that owns none)."""
    if not params:
        return {}
    M = num_instances or max(len(np.atleast_1d(v)) for v in params.values())
    idx = shard_indices(M, rank, world)
    per = shard_size(M, world)
    if len(idx) < per:
        fill = idx[-1] if len(idx) else 0
        idx = np.concatenate([idx, np.full(per - len(idx), fill, dtype=idx.dtype)])
    return {k: np.ascontiguousarray(np.broadcast_to(np.asarray(v), (M,))[idx])
            for k, v in params.items()}


def interleave(blocks, num_instances):
    """blocks[r] = [per, ...] of rank r -> [M, ...] in global instance order
    (global g = local * W + rank); padding falls off the end."""
    W = len(blocks)
    per = blocks[0].shape[0]
    out = np.empty((per * W,) + blocks[0].shape[1:], dtype=blocks[0].dtype)
    for r, b in enumerate(blocks):
        out[r::W] = b
    return out[:num_instances]


def _device_of(group):
    import torch
    import torch.distributed as dist
    if dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_array(arr, src=0, group=None):
    """Broadcast a small numpy array (shape and dtype known on every rank)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(arr).copy()).to(_device_of(group))
    dist.broadcast(t, src=src, group=group)
    return t.cpu().numpy()


class Gatherer:
    """The waveform gather of one process group.  On NCCL it owns a
    simcuComm_ of the C-ABI (its own NCCL communicator over the same ranks)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nccl = dist.get_backend(group) == "nccl"
        self.comm = None
        self.last = None          # simcuGatherStats_ of the last gather, as a dict
        if self.nccl:
            from . import engine
            L = engine.lib()
            ident = np.zeros(128, dtype=np.uint8)
            if self.rank == 0 and L.simcuCommUniqueId(ident.ctypes.data, 128):
                raise RuntimeError("simcuCommUniqueId failed")
            ident = broadcast_array(ident, 0, group)
            self.L = L
            self.comm = C.c_void_p()
            if L.simcuCommInit(C.byref(self.comm), self.world, self.rank, ident.ctypes.data, 128):
                raise RuntimeError("simcuCommInit failed")

    def gather(self, eng, num_instances, root=0, out=None, status=None):
        """All ranks call this after ``eng.run_device()``.  The root returns
        (data [M, ngrid, nprobes], status [M]) in global instance order, the
        others (None, None).  ``out`` / ``status`` (root): preallocated host
        arrays for W * shard_size instances (e.g. pinned), else numpy allocates."""
        per = eng.num_instances
        W = self.world
        nprobes = len(eng._run[2])
        shape = (per * W, eng.ngrid, nprobes)
        if self.nccl:
            from . import engine
            st = engine.GatherStats()
            if self.rank == root:
                if out is None:
                    out = np.empty(shape, dtype=np.float64)
                if status is None:
                    status = np.empty(per * W, dtype=np.int32)
                rc = self.L.simcuGatherOutput(eng.h, self.comm, root, out.ctypes.data,
                                              status.ctypes.data, None, C.byref(st))
            else:
                rc = self.L.simcuGatherOutput(eng.h, self.comm, root, None, None, None, C.byref(st))
            if rc:
                raise RuntimeError("simcuGatherOutput failed")
            self.last = st.as_dict()
            if self.rank != root:
                return None, None
            return out[:num_instances], status[:num_instances]
        # gloo: host blocks (CPU tests of the sharding logic)
        data, stat = eng.fetch()
        return self.gather_host(data, stat, num_instances, root)

    def gather_async(self, eng, root=0, out=None, status=None):
        """NCCL only: queue the gather of the batch `eng` has just run and return at
        once, so the next batch can be launched while the waveforms travel.  The
        root passes its (pinned) host arrays for W * shard_size instances; they are
        complete after ``wait()``."""
        if not self.nccl:
            raise RuntimeError("gather_async needs the NCCL backend")
        if self.rank == root:
            rc = self.L.simcuGatherOutputAsync(eng.h, self.comm, root, out.ctypes.data,
                                               status.ctypes.data, None)
        else:
            rc = self.L.simcuGatherOutputAsync(eng.h, self.comm, root, None, None, None)
        if rc:
            raise RuntimeError("simcuGatherOutputAsync failed")

    def wait(self):
        """Completes the gather queued by gather_async (no-op if none is pending)."""
        from . import engine
        st = engine.GatherStats()
        if self.L.simcuGatherWait(self.comm, C.byref(st)):
            raise RuntimeError("simcuGatherWait failed")
        self.last = st.as_dict()
        return self.last

    def gather_host(self, data, stat, num_instances, root=0):
        """gloo path: dist.gather of [per, ...] host blocks, interleaved on the root."""
        import torch
        import torch.distributed as dist
        td = torch.from_numpy(np.ascontiguousarray(data))
        ts = torch.from_numpy(np.ascontiguousarray(stat))
        if self.rank == root:
            ld = [torch.empty_like(td) for _ in range(self.world)]
            ls = [torch.empty_like(ts) for _ in range(self.world)]
            dist.gather(td, ld, dst=root, group=self.group)
            dist.gather(ts, ls, dst=root, group=self.group)
            return (interleave([x.numpy() for x in ld], num_instances),
                    interleave([x.numpy() for x in ls], num_instances))
        dist.gather(td, None, dst=root, group=self.group)
        dist.gather(ts, None, dst=root, group=self.group)
        return None, None

    def close(self):
        if self.comm is not None and self.comm:
            self.L.simcuCommDestroy(C.byref(self.comm))
            self.comm = None


def tran_batch_sharded(circuit, tstep, tstop, tmax=0.0, params=None, probes=(), group=None,
                       root=0, gatherer=None, **options):
    """``Circuit.tran_batch`` over all ranks of the process group: each rank
    simulates its round-robin shard on its current CUDA device, the waveforms
    are gathered to ``root``.  The root returns a BatchResult of the whole sweep
    (global instance order, summed statistics); the other ranks return one with
    ``data = status = None``.  All ranks use the pivot sequences the root derived
    on its pilots, so the result does not depend on the number of ranks."""
    import torch
    import torch.distributed as dist
    from . import engine, units
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    params = params or {}
    M = max(len(np.atleast_1d(v)) for v in params.values()) if params else 1
    mine = shard_params(params, rank, world, M)
    per = shard_size(M, world)
    tstep, tstop, tmax = units.float(tstep), units.float(tstop), units.float(tmax)
    own = gatherer is None
    g = gatherer or Gatherer(group)
    eng = engine.Engine(circuit.netlist(), per)
    for k, v in options.items():
        eng.set_option(k, v)
    eng.set_params(mine)
    n = eng.num_unknowns()
    if rank == root:
        eng.prepare(tstep, tstop, tmax, list(probes))
        piv = np.stack(eng.pivots(0) + eng.pivots(1)).astype(np.int32)
    else:
        piv = np.zeros((4, n), dtype=np.int32)
    piv = broadcast_array(piv, root, group)
    if rank != root:
        eng.set_pivots(0, piv[0], piv[1])
        eng.set_pivots(1, piv[2], piv[3])
        eng.prepare(tstep, tstop, tmax, list(probes))
    eng.run_device()
    data, status = g.gather(eng, M, root)
    # statistics: padded instances are simulated but not counted
    own_real = len(shard_indices(M, rank, world))
    acc = eng.counters("accepted")[:own_real].sum()
    stat_local = eng.counters("status")[:own_real]
    st = eng.stats()
    dev = _device_of(group)
    counters = torch.tensor([float(acc), float((stat_local != 0).sum()),
                             float(eng.counters("rejected_lte")[:own_real].sum()),
                             float(eng.counters("rejected_nc")[:own_real].sum()),
                             float(eng.counters("factorizations")[:own_real].sum())],
                            dtype=torch.float64, device=dev)
    dist.all_reduce(counters, group=group)
    ms = torch.tensor([st["deviceMs"]], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=group)
    total = dict(zip(("accepted", "failed", "rejectedLTE", "rejectedNC", "factorizations"),
                     [int(x) for x in counters.tolist()]))
    total["deviceMs"] = float(ms.item())
    total["gather"] = g.last
    ngrid = eng.ngrid
    eng.close()
    if own:
        g.close()
    return engine.BatchResult(data, engine.output_grid(tstep, tstop, ngrid), status, total,
                              list(probes))
