#!/usr/bin/env python
"""Adds tests/golden/ibis_vt.npz: the V-t waveform tables, fixtures and ramp
data of the reference's IBIS fixture (module/ibis_test.py) as parsed by the
UNMODIFIED reference parser -- the inputs of its switching-coefficient (K) table
generation (module/ibis_driver.py).  The K tables themselves are already in
ibis_tables.npz (made by make_golden.py).  Build container only.

Usage:  python tests/golden/make_ibis_vt.py
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402

CODE = r'''
import sys, json, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, %r)
import numpy as np
import eispice
ibs = eispice.Ibis("test")
arrays, meta = {}, {}
for name, m in ibs.model.items():
    mm = {"waveforms": {}, "ramp": None}
    for direction in ("rising", "falling"):
        waves = getattr(m, direction + "_waveform", None)
        if waves is None:
            continue
        mm["waveforms"][direction] = []
        for k, w in enumerate(waves):
            ent = {}
            for key in ("r_fixture", "v_fixture", "v_fixture_min", "v_fixture_max",
                        "c_fixture", "l_fixture"):
                if hasattr(w, key):
                    ent[key] = float(getattr(w, key))
            for sp in ("typ", "min", "max"):
                arrays["%%s/%%s/%%d/%%s" %% (name, direction, k, sp)] = np.asarray(w.data[sp], dtype=np.float64)
            mm["waveforms"][direction].append(ent)
    if hasattr(m, "ramp"):
        r = m.ramp
        mm["ramp"] = {"r_load": float(r.r_load)}
        for key in ("dv_r", "dt_r", "dv_f", "dt_f"):
            if hasattr(r, key):
                mm["ramp"][key] = {sp: float(v) for sp, v in getattr(r, key).items()}
    meta[name] = mm
np.savez_compressed(%r, meta=json.dumps(meta), **arrays)
print("ibis V-t:", len(arrays), "arrays;", {k: {d: len(v) for d, v in mm["waveforms"].items()} for k, mm in meta.items()})
'''

if __name__ == "__main__":
    mod = make_golden.build_reference_python()
    subprocess.run([sys.executable, "-c", CODE % (mod, os.path.join(HERE, "ibis_vt.npz"))],
                   check=True, cwd=make_golden.TMP)
