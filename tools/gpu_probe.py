#!/usr/bin/env python
"""Development probe (not a test): runs every golden circuit through the C-ABI
on the GPU, compares with the oracle replaying the GPU's pivot sequence and
with the committed reference waveforms, and prints one line per circuit."""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import backends  # noqa: E402
import circuits  # noqa: E402
import parity  # noqa: E402
from eispice_b200 import engine, units  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ref_waveforms.npz"))
META = json.loads(str(GOLD["meta"]))


def main():
    backends.build_oracle()
    only = sys.argv[1:]
    report = {}
    dump = {}
    for name, (build, an, checks) in sorted(circuits.all_cases().items()):
        if only and name not in only:
            continue
        t0 = time.time()
        try:
            nl = build().netlist()
            eng = engine.Engine(nl, 1)
            if an[0] == "op":
                data, var = eng.op_single()
            else:
                data, var = eng.tran_single(*[units.float(x) for x in an[1:]])
            st = eng.stats()
            ora = backends.OracleSim(nl)
            pc0, pr0 = eng.pivots(0)
            ora.set_pivots(0, pc0, pr0)
            if an[0] != "op":
                pc1, pr1 = eng.pivots(1)
                ora.set_pivots(1, pc1, pr1)
                od, ov = ora.tran(*[units.float(x) for x in an[1:]])
            else:
                od, ov = ora.op()
            line = {"npts": int(data.shape[0]), "npts_oracle": int(od.shape[0]),
                    "N": st["numUnknowns"], "nnzA": st["numNonzeros"],
                    "F": st["numFactor"], "Fop": st["numFactorOP"],
                    "fact": st["factorizations"], "rejL": st["rejectedLTE"],
                    "rejN": st["rejectedNC"], "ms": round(st["deviceMs"], 2)}
            assert var == ov
            if data.shape == od.shape:
                frac, worst = parity.tight_fraction(data, od)
                line["tight_frac"] = round(frac, 5)
                line["tight_worst"] = float("%.3g" % worst)
            if an[0] != "op":
                tstep, tstop = units.float(an[1]), units.float(an[2])
                line["grid_vs_oracle"] = float("%.3g" % parity.grid_error(
                    data[:, 0], data, od[:, 0], od, var, tstep, tstop))
                ref = GOLD[name]
                rvar = META[name]["variables"]
                idx = [var.index(v) for v in rvar]
                line["grid_vs_ref"] = float("%.3g" % parity.grid_error(
                    data[:, 0], data[:, idx], ref[:, 0], ref, rvar, tstep, tstop))
            dump[name] = data
            dump[name + "__oracle"] = od
            ok = True
            for kind, node, value, tm in checks:
                col = var.index("%s(%s)" % (kind, node))
                got = (float(np.interp(units.float(tm), data[:, 0], data[:, col]))
                       if len(data) > 1 else float(data[0, col]))
                ok = ok and parity.reference_check(units.float(value), got)
            line["known_answers"] = ok
            eng.close()
        except Exception as e:  # noqa: BLE001
            line = {"error": "%s: %s" % (type(e).__name__, e)}
            traceback.print_exc()
        line["wall_s"] = round(time.time() - t0, 2)
        report[name] = line
        print(name, json.dumps(line), flush=True)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(report, open(os.path.join(out, "probe.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(out, "probe_data.npz"), **dump)


if __name__ == "__main__":
    main()
