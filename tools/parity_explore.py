#!/usr/bin/env python
"""GPU-box exploration behind the parity tests: for a random sample of a
full-size sweep, (1) adaptive GPU run vs oracle / reference on the tstep grid,
(2) replay of the oracle's attempt log on the GPU vs the oracle at 1e-6/1e-9.
usage: parity_explore.py cfg [nsample] [out.json]"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import backends  # noqa: E402
import parity  # noqa: E402
import parity_workers as pw  # noqa: E402
import workloads  # noqa: E402
from eispice_b200 import engine  # noqa: E402

CAP = {"cfg1": 4000, "cfg2": 6000, "cfg3": 4000, "cfg4": 16000, "cfg5": 40000}


def main():
    cfg = sys.argv[1]
    ns = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    out = sys.argv[3] if len(sys.argv) > 3 else None
    w = workloads.make(cfg)
    rng = np.random.default_rng(99)
    pick = np.sort(rng.choice(w.M, min(ns, w.M), replace=False))
    sub = {k: np.asarray(v)[pick] for k, v in w.params.items()}
    n = len(pick)
    cap = CAP[cfg]
    backends.build_oracle()
    pool = mp.get_context("spawn").Pool(len(os.sched_getaffinity(0)))
    rep = {"cfg": cfg, "M_full": w.M, "sample": n}

    eng = engine.Engine(w.cct.netlist(), n)
    t0 = time.time()
    res = eng.tran_batch(w.tstep, w.tstop, 0.0, sub, w.probes, raw_max_points=cap, raw_count=n)
    rep["gpu_adaptive_s"] = time.time() - t0
    rep["status"] = np.bincount(res.status).tolist()
    raws = eng.fetch_raw_all(cap)
    names = eng.variables()[1:]
    piv = eng.pivots(0) + eng.pivots(1)

    t0 = time.time()
    logged = pool.map(pw.oracle_logged, [(cfg, w.M, int(i), piv) for i in pick])
    nat = pool.map(pw.gridded, [(cfg, w.M, int(i), "oracle") for i in pick])
    ref = pool.map(pw.gridded, [(cfg, w.M, int(i), "ref") for i in pick]) if backends.ref_available() else [None] * n
    rep["cpu_s"] = time.time() - t0

    # (1) adaptive: GPU vs natural oracle vs reference on the grid
    e_gpu, e_ref, e_gpu_f, e_ref_f, e_gpu_ref = [], [], [], [], []
    for j in range(n):
        if res.status[j] != 0 or nat[j] is None:
            continue
        g = pw.to_grid(raws[j], w.tstep, w.tstop)
        e_gpu.append(pw.spice_ratio(g, nat[j], names))
        e_gpu_f.append(pw.spice_ratio(g, nat[j], names, 1e-3))
        if ref[j] is not None:
            e_ref.append(pw.spice_ratio(ref[j], nat[j], names))
            e_ref_f.append(pw.spice_ratio(ref[j], nat[j], names, 1e-3))
            e_gpu_ref.append(pw.spice_ratio(g, ref[j], names))

    def summ(x):
        x = np.array(x)
        return {"n": len(x), "frac_le_1": float((x <= 1).mean()), "median": float(np.median(x)),
                "p90": float(np.quantile(x, 0.9)), "max": float(x.max())} if len(x) else {}
    rep["adaptive"] = {"gpu_vs_oracle": summ(e_gpu), "ref_vs_oracle": summ(e_ref),
                       "gpu_vs_ref": summ(e_gpu_ref),
                       "gpu_vs_oracle_floor": summ(e_gpu_f), "ref_vs_oracle_floor": summ(e_ref_f)}

    # same-sequence count of the adaptive run
    same = 0
    for j in range(n):
        if logged[j] is not None and raws[j].shape == logged[j][0].shape:
            fr, worst = parity.tight_fraction(raws[j], logged[j][0])
            same += worst <= 1.0
    rep["adaptive_same_sequence_tight"] = same

    # (2) replay
    ok = [j for j in range(n) if logged[j] is not None]
    logs = [(logged[j][1], logged[j][2], logged[j][3]) if logged[j] is not None
            else (np.zeros(0), np.zeros(0, np.int32), np.zeros(0, np.int32)) for j in range(n)]
    eng.set_replay_log(logs)
    eng.set_option("fmad", float(os.environ.get("EXPLORE_FMAD", "0")))
    t0 = time.time()
    res2 = eng.tran_batch(w.tstep, w.tstop, 0.0, sub, w.probes, raw_max_points=cap, raw_count=n)
    rep["gpu_replay_s"] = time.time() - t0
    raws2 = eng.fetch_raw_all(cap)
    diffs = eng.counters("replay_diff")
    ldiffs = eng.counters("linearize_diff")
    worst_all, npass, nshape, bad = 0.0, 0, 0, []
    total = inside = 0
    for j in ok:
        od = logged[j][0]
        if raws2[j].shape != od.shape:
            nshape += 1
            bad.append((int(pick[j]), "shape", raws2[j].shape[0], od.shape[0]))
            continue
        err = np.abs(raws2[j] - od)
        tol = parity.RTOL_TIGHT * np.abs(od) + parity.ATOL_TIGHT
        r = err / tol
        total += r.size
        inside += int((r <= 1).sum())
        wj = float(r.max())
        worst_all = max(worst_all, wj)
        if wj <= 1.0:
            npass += 1
        else:
            k = np.unravel_index(np.argmax(r), r.shape)
            bad.append((int(pick[j]), "value", wj, int(k[0]), names[k[1] - 1] if k[1] else "time",
                        float(od[k]), float(raws2[j][k]), int(diffs[j]), int(ldiffs[j])))
    rep["replay"] = {"instances": len(ok), "pass_all_samples": npass, "shape_mismatch": nshape,
                     "worst_ratio": worst_all, "samples": total, "samples_inside": inside,
                     "replay_diffs_total": int(diffs.sum()), "instances_with_diffs": int((diffs > 0).sum()),
                     "linearize_diffs_total": int(ldiffs.sum()), "instances_with_lin_diffs": int((ldiffs > 0).sum()),
                     "linearize_passes": int(sum(len(x[2]) for x in logs)),
                     "status": np.bincount(res2.status).tolist(), "bad": bad[:20]}
    print(json.dumps(rep, indent=1))
    if out:
        json.dump(rep, open(out, "w"), indent=1)
    pool.close()


if __name__ == "__main__":
    main()
