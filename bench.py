#!/usr/bin/env python
"""bench.py -- instance-timesteps/s of the batched transient engine on B200.

One "step" = one batched transient (load -> operating point -> every timestep
of every instance) over one batch of synthetic parameter sets of the workload.
Default workload: BASELINE.json configs[4], the north-star run - IBIS driver +
8 lossy T-line sections + receiver Monte-Carlo, 131 072 instances PER GPU (weak
scaling: 8 GPUs = the 1 048 576-instance sweep; every rank simulates its own
shard with its own seed; instances are independent, so there is no data-path
collective - the one NCCL gather of the waveforms to rank 0 is part of e2e).

  value     accepted instance-timesteps of all ranks / device time, inputs
            already resident in HBM (CUDA events on the launching stream, max
            over ranks)
  e2e       the same through the reference-facing C-ABI calls with HOST
            buffers: per-instance parameters copied host->device and the
            gridded probe waveforms + status copied device->host every step;
            at N > 1 the waveforms of all ranks are gathered to rank 0 over
            NCCL (device buffers) and copied to ITS host inside the timer
  roofline  transient kernel: SURVEY.md 8d algorithmic bytes per accepted
            instance-timestep x timesteps / kernel time, vs measured HBM peak
  cpu_baseline  the unmodified reference (oracle/_ref) on the host cores, one
            process per core, on a bounded sample of the same sweep

`--impl reference` times only that CPU reference arm.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "instance-timesteps/sec batched transient"
UNIT = "instance-timesteps/s"

_JSON_OUT = None


def _claim_stdout():
    """stdout must carry the one JSON line only: libraries that print to fd 1
    (NCCL's version banner, simcuInfo, ...) are sent to stderr from here on."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


# ---------------------------------------------------------------------------
# CPU reference arm (test infrastructure: oracle/_ref, else the oracle port)
# ---------------------------------------------------------------------------
def _cpu_worker(args):
    name, M, indices, budget_s, use_ref = args
    import backends
    import workloads
    w = workloads.make(name, M)
    steps = n = aborted = 0
    t0 = time.perf_counter()
    for i in indices:
        nl, s = w.instance_netlist(int(i))
        sim = backends.RefSim(nl, s) if use_ref else backends.OracleSim(nl, s)
        try:
            data, _ = sim.tran(w.tstep, w.tstop)
            steps += data.shape[0] - 1
        except RuntimeError:
            # the reference aborts the run ("Timestep too small", simulator.c:193-195);
            # its time is spent, no timestep is credited
            aborted += 1
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return steps, n, time.perf_counter() - t0, aborted


def cpu_reference_rate(name, M, budget_s, pool=None, cores=None, seed=7):
    """(instance-timesteps/s, cores, kind, sample text, wall seconds) of the CPU arm."""
    import backends
    backends.build_oracle()
    use_ref = backends.ref_available()
    cores = cores or len(os.sched_getaffinity(0))
    rng = np.random.default_rng(seed)
    order = rng.permutation(M)
    chunks = [order[c::cores][:4096] for c in range(cores)]
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(cores)
    try:
        res = pool.map(_cpu_worker, [(name, M, ch, budget_s, use_ref) for ch in chunks])
    finally:
        if own:
            pool.close()
    steps = sum(r[0] for r in res)
    n = sum(r[1] for r in res)
    wall = max(r[2] for r in res)
    aborted = sum(r[3] for r in res)
    kind = "reference" if use_ref else "port"
    sample = "%d of %d instances of %s (random subset), %d processes, %.1f s" % (
        n, M, name, cores, wall)
    if aborted:
        sample += ", %d aborted by the reference (timestep too small)" % aborted
    return steps / wall, cores, kind, sample, wall


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = a.workload
    import workloads
    M = a.instances or workloads.DEFAULT_INSTANCES[name]
    cores = len(os.sched_getaffinity(0))
    pool = mp.get_context("spawn").Pool(cores)
    rates, walls = [], []
    sample = kind = None
    # every step is a fresh random sample of the sweep, time-budgeted so that the
    # whole --steps/--warmup run stays within a few minutes
    budget = min(a.cpu_seconds, max(4.0, 240.0 / max(1, a.warmup + a.steps)))
    try:
        for it in range(a.warmup + a.steps):
            rate, cores, kind, sample, wall = cpu_reference_rate(name, M, budget, pool, cores,
                                                                 seed=7 + it)
            if it >= a.warmup:
                rates.append(rate)
                walls.append(wall)
    finally:
        pool.close()
    value = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * float(np.mean(walls)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s" % (name, workloads.DESCRIPTION[name]),
                       "instances_per_gpu": M},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "per step (fresh random subset each): " + sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    _emit(line)


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def other_workloads(peak):
    """Informational: the other BASELINE configurations at their named sizes -
    cfg2 @ 10 000 (round 1's headline) and cfg4 @ 262 144 (the largest single-GPU
    config) - one warm-up + one timed batch each, device-resident timing only."""
    import workloads
    from eispice_b200 import engine
    out = []
    for name, M in (("cfg2", 10000), ("cfg4", 262144)):
        try:
            w = workloads.make(name, M)
            eng = engine.Engine(w.cct.netlist(), w.M)
            eng.set_params(w.params)
            eng.prepare(w.tstep, w.tstop, 0.0, w.probes)
            eng.run_device()
            eng.run_device()
            st = eng.stats()
            n_tlines = sum(1 for d in w.cct.netlist() if d[0] == "T")
            balg, k = workloads.algorithmic_bytes(st, n_tlines, len(w.probes))
            rate = st["accepted"] / (st["deviceMs"] * 1e-3)
            out.append({"workload": "%s: %s" % (name, workloads.DESCRIPTION[name]),
                        "instances": M, "value": rate, "unit": UNIT,
                        "ms_per_step": st["deviceMs"], "solves_per_timestep": round(k, 3),
                        "bytes_per_instance_timestep": round(balg, 1),
                        "roofline_frac": balg * rate / 1e9 / peak,
                        "failed_instances": st["failed"],
                        "registers_per_thread": st["registersPerThread"],
                        "local_bytes_per_thread": st["localBytesPerThread"]})
            eng.close()
        except Exception as e:  # noqa: BLE001
            out.append({"workload": name, "error": str(e)})
    return out


def _pinned(shape, dtype):
    """Pinned host array (falls back to pageable memory if the box refuses)."""
    import torch
    try:
        return torch.empty(shape, dtype=dtype, pin_memory=True).numpy(), True
    except RuntimeError:
        return torch.empty(shape, dtype=dtype).numpy(), False


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--instances", type=int, default=0, help="instances per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=30.0)
    ap.add_argument("--e2e-steps", type=int, default=3, help="cap on timed end-to-end batches")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the one-batch timing of cfg2@10k and cfg4@262144 (N=1 only)")
    ap.add_argument("--tpb", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--minblocks", type=int, default=0)
    ap.add_argument("--maxblocks", type=int, default=0)
    ap.add_argument("--claim-below", type=int, default=-1)
    a = ap.parse_args()
    _claim_stdout()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup

    if a.impl == "reference":
        run_reference_arm(a)
        return

    import torch
    import torch.distributed as dist
    import workloads
    from eispice_b200 import engine, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local)
    engine.set_device(local)
    gatherer = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        gatherer = sharding.Gatherer()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    name = a.workload
    M = a.instances or workloads.DEFAULT_INSTANCES[name]
    w = workloads.make(name, M, seed_offset=1000 * rank)
    eng = engine.Engine(w.cct.netlist(), w.M)
    if a.tpb:
        eng.set_option("threads_per_block", a.tpb)
    if a.lanes:
        eng.set_option("lanes_per_warp", a.lanes)
    if a.minblocks:
        eng.set_option("min_blocks_per_sm", a.minblocks)
    if a.maxblocks:
        eng.set_option("max_blocks_per_sm", a.maxblocks)
    if a.claim_below >= 0:
        eng.set_option("claim_below", a.claim_below)
    eng.set_params(w.params)
    t0 = time.perf_counter()
    eng.prepare(w.tstep, w.tstop, 0.0, w.probes)
    prepare_s = time.perf_counter() - t0

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    stream = torch.cuda.current_stream()

    # -- device-resident timing ------------------------------------------------
    for _ in range(a.warmup):
        eng.run_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms = tran_ms = 0.0
    launches = 0
    accepted = 0
    for _ in range(a.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        eng.run_device()                       # legacy default stream == torch's default
        e1.record(stream)
        e1.synchronize()
        dev_ms += e0.elapsed_time(e1)
        st = eng.stats()
        tran_ms += st["tranMs"]
        launches += st["kernelLaunches"]
        accepted += st["accepted"]
    barrier()
    st = eng.stats()
    failed = st["failed"]

    # -- end to end: host buffers in, host buffers out ---------------------------
    nprobes = len(w.probes)
    total = w.M * world
    pinned = True
    if rank == 0:
        out, pinned = _pinned((total, eng.ngrid, nprobes), torch.float64)
        status, _ = _pinned((total,), torch.int32)
    else:
        out = status = None
    h2d = sum(np.asarray(v).size * (4 if k.endswith(".set") else 8) for k, v in w.params.items())
    d2h = (out.nbytes + status.nbytes) if rank == 0 else 0
    gather = {"ms": 0.0, "copy_ms": 0.0, "bytes": 0}

    def e2e_step():
        eng.set_params(w.params)               # host arrays -> the handle (as the binding does)
        eng.upload()                           # host -> device
        eng.run_device()
        if world == 1:
            eng.fetch(out, status)             # device -> host, [M][ngrid][nprobes] + status
            return 1
        # all ranks -> rank 0 over NCCL (device buffers), then ONE strided copy per rank
        # to rank 0's host.  Pipelined: the transfer of this batch runs on the
        # communicator's stream while the next batch computes; it is waited for
        # before the next gather and at the end of the timed region.
        e2e_wait()
        gatherer.gather_async(eng, 0, out, status)
        return 2 + (world if rank == 0 else 0)   # transpose + gather copies

    def e2e_wait():
        if world == 1:
            return
        g = gatherer.wait()
        if g and g["ncclBytes"]:
            gather["ms"] += g["ncclMs"]
            gather["copy_ms"] += g["copyMs"]
            gather["bytes"] += g["ncclBytes"]
            gather["count"] = gather.get("count", 0) + 1

    e2e_step()
    e2e_wait()
    barrier()
    for key in list(gather):
        gather[key] = 0
    e2e_steps = max(1, min(a.steps, a.e2e_steps))
    t0 = time.perf_counter()
    e2e_accepted = 0
    for _ in range(e2e_steps):
        extra = e2e_step()
        e2e_accepted += eng.stats()["accepted"]
        launches += eng.stats()["kernelLaunches"] + extra
    e2e_wait()                                 # the last batch's waveforms are on the host
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_s = time.perf_counter() - t0
    codes, counts = np.unique(eng.counters("status"), return_counts=True)
    status_hist = {engine.STATUS.get(int(c), str(int(c))): int(n) for c, n in zip(codes, counts)}
    clocks = sampler.stop() if rank == 0 else None
    barrier()

    # -- reduce over ranks -------------------------------------------------------
    vals = torch.tensor([dev_ms, e2e_s, tran_ms], dtype=torch.float64, device="cuda")
    sums = torch.tensor([accepted, e2e_accepted, failed, launches, h2d, d2h], dtype=torch.float64,
                        device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_s_max, tran_ms_max = vals.tolist()
    accepted_all, e2e_accepted_all, failed_all, launches_all, h2d_all, d2h_all = sums.tolist()

    if rank == 0:
        value = accepted_all / (dev_ms_max * 1e-3)
        e2e_value = e2e_accepted_all / e2e_s_max
        n_tlines = sum(1 for d in w.cct.netlist() if d[0] == "T")
        balg, k = workloads.algorithmic_bytes(st, n_tlines, nprobes)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # per launch: this rank's accepted steps of ONE run x B_alg / tran kernel time
        steps_per_launch = accepted / a.steps
        achieved = balg * steps_per_launch / (tran_ms / a.steps * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": dev_ms_max / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s" % (name, workloads.DESCRIPTION[name]),
                       "instances_per_gpu": w.M, "instances_total": total,
                       "unknowns": st["numUnknowns"],
                       "nnzA": st["numNonzeros"], "nnzLU": st["numFactor"],
                       "solves_per_timestep": round(k, 3),
                       "timesteps_per_instance": round(accepted / a.steps / w.M, 1),
                       "threads_per_block": st["threadsPerBlock"],
                       "grid_blocks": st["gridBlocks"], "lanes_per_warp": st["lanesPerWarp"],
                       "registers_per_thread": st["registersPerThread"],
                       "local_bytes_per_thread": st["localBytesPerThread"],
                       "l2": "flushed between timed steps (256 MiB write)",
                       "failed_instances": int(failed_all), "status_rank0": status_hist,
                       "prepare_s": round(prepare_s, 3),
                       "nvrtc_s": round(st["jitCompileMs"] * 1e-3, 3), "nvrtc": engine.nvrtc_path(),
                       "e2e_steps": e2e_steps,
                       "e2e_host_buffers": "pinned" if pinned else "pageable"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all),
                    "d2h_bytes_per_step": int(d2h_all)},
            "gather": None if world == 1 else {
                "to_rank": 0, "bytes": int(gather["bytes"] / max(1, gather.get("count", 1))),
                "ms": gather["ms"] / max(1, gather.get("count", 1)),
                "host_copy_ms": gather["copy_ms"] / max(1, gather.get("count", 1)),
                "how": "grouped ncclSend/ncclRecv between device buffers (simcuGatherOutputAsync), "
                       "inside the e2e timer, overlapped with the next batch; ms includes waiting "
                       "for the slowest rank's kernel"},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None,
                         "kernel": "simcuJitKernel",
                         "bytes_per_instance_timestep": round(balg, 1),
                         "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
            "clocks": clocks,
        }
        # measured DRAM traffic of the dominant kernel, per launch, from the committed
        # ncu capture of this same workload (profiles/summary.json), else null; next
        # to it dram_frac = measured DRAM bytes/s / measured HBM peak and issue_frac =
        # issue slots used / available (148 SMs x 4 schedulers), both from that
        # capture: what actually bounds the kernel is the latency of its thread-local
        # state streaming through L2/HBM, not the algorithmic byte model.
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "summary.json")))
            ent = prof.get("%s@%d" % (name, w.M))
            if ent:
                line["roofline"]["traffic"] = ent["dram_bytes_per_launch"]
                line["roofline"]["traffic_source"] = ent["source"]
                if ent.get("warp_instructions_per_instance_timestep"):
                    ipt = ent["warp_instructions_per_instance_timestep"]
                    mhz = (clocks or {}).get("sm_mhz") or 1965.0
                    line["roofline"]["issue_frac"] = ipt * (value / world) / (148 * 4 * mhz * 1e6)
                    line["roofline"]["warp_instructions_per_instance_timestep"] = ipt
                elif ent.get("issue_active_frac"):
                    # measured directly by ncu on that capture (sm__issue_active)
                    line["roofline"]["issue_frac"] = ent["issue_active_frac"]
                if ent.get("dram_throughput_frac_of_measured_peak"):
                    line["roofline"]["dram_frac"] = ent["dram_throughput_frac_of_measured_peak"]
        except (OSError, ValueError, KeyError):
            pass
        if world == 1 and not a.no_extras:
            eng.close()
            eng = None
            del flush
            torch.cuda.empty_cache()
            line["other_workloads"] = other_workloads(peak)
        if world == 1 and not a.no_cpu_baseline:
            try:
                rate, cores, kind, sample, _ = cpu_reference_rate(name, w.M, a.cpu_seconds)
                line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores,
                                        "kind": kind, "sample": sample}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0,
                                        "kind": "port", "sample": "failed: %s" % e}
        _emit(line)
    if eng is not None:
        eng.close()
    if gatherer is not None:
        gatherer.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
