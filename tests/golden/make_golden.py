#!/usr/bin/env python
"""Generates the committed golden fixtures from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference); the fixtures travel,
the reference does not.

  A. tests/golden/ibis_tables.npz -- V-I and switching-coefficient (K) tables of
     the reference's own IBIS fixture (module/ibis_test.py), produced by the
     reference's parser + K-table generator (module/ibis_parser.py,
     module/ibis_driver.py) running on the reference engine.  For this the
     reference's CPython binding is compiled under /tmp from a copy with the
     compatibility edits of SURVEY.md 8c (two "u" format units -> "s", numpy-2
     macro names, a tkinter-free plot stub, one regex flag position); no
     arithmetic is touched.
  B. tests/golden/ref_waveforms.npz -- result arrays of the reference engine
     (oracle/_ref/libeispice_ref.so, driven through ctypes) for every circuit
     in tests/circuits.py, at the reference's own adaptive time points.

Usage:  python tests/golden/make_golden.py
"""
import glob
import json
import os
import shutil
import subprocess
import sys
import sysconfig
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("REF", "/root/reference")
TMP = "/tmp/eisref"


def build_reference_python():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    mod = os.path.join(TMP, "mod")
    shutil.rmtree(TMP, ignore_errors=True)
    os.makedirs(mod)
    for f in glob.glob(os.path.join(REF, "module", "*.py")):
        shutil.copy(f, mod)
    src = open(os.path.join(REF, "module", "simulatormodule.c")).read()
    src = src.replace('"OOuU:B"', '"OOsU:B"').replace('"OOuOO:CB"', '"OOsOO:CB"')
    open(os.path.join(TMP, "simulatormodule.c"), "w").write(src)
    p = os.path.join(mod, "ibis_parser.py")
    s = open(p).read().replace('self.addRE("\\s*" + _reRange % "\\S*", self.handle)',
                               'self.addRE(_reRange % "\\S*", self.handle)')
    open(p, "w").write(s)
    open(os.path.join(mod, "plot.py"), "w").write(
        "def plot(*a, **k): pass\ndef plot_voltage(*a, **k): pass\n"
        "def plot_current(*a, **k): pass\n")
    inc = []
    for d in ("log", "data", "calculon", "superlu", "simulator/include"):
        inc += ["-I", os.path.join(REF, "libs", d)]
    inc += ["-I", sysconfig.get_paths()["include"], "-I", np.get_include()]
    cc = ["gcc", "-O2", "-fPIC", "-w", "-std=gnu99"]
    subprocess.run(cc + inc + ["-DPyArray_DOUBLE=NPY_DOUBLE",
                               "-DNPY_OWNDATA=NPY_ARRAY_OWNDATA", "-c",
                               os.path.join(TMP, "simulatormodule.c"), "-o",
                               os.path.join(TMP, "simulatormodule.o")], check=True)
    subprocess.run(cc + ["-DREF_SHIM_NO_LOG_STREAMS", "-c",
                         os.path.join(ROOT, "oracle", "ref_shim.c"), "-o",
                         os.path.join(TMP, "shim_nolog.o")], check=True)
    objs = [o for o in glob.glob(os.path.join(ROOT, "oracle", "_ref", "obj", "*.o"))
            if not o.endswith("ref_shim.o")]
    subprocess.run(["gcc", "-shared", "-o", os.path.join(mod, "simulator_.so"),
                    os.path.join(TMP, "simulatormodule.o"),
                    os.path.join(TMP, "shim_nolog.o")] + objs + ["-lm"], check=True)
    return mod


def dump_ibis_tables(mod):
    code = r'''
import sys, json, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, %r)
import numpy as np
import eispice
ibs = eispice.Ibis("test")
comp = ibs.component[ibs.device]
arrays, meta = {}, {"models": {}, "pins": {}, "package": {}}
for name, m in ibs.model.items():
    mm = {"model_type": m.model_type, "voltage_range": m.voltage_range,
          "c_comp": m.c_comp, "tables": [], "k_tables": []}
    for tab in ("gnd_clamp", "power_clamp", "pullup", "pulldown"):
        if hasattr(m, tab):
            mm["tables"].append(tab)
            for sp, arr in getattr(m, tab).items():
                arrays["%%s/%%s/%%s" %% (name, tab, sp)] = np.asarray(arr, dtype=np.float64)
    for tab in ("pullup_k", "pulldown_k"):
        if hasattr(m, tab):
            mm["k_tables"].append(tab)
            for dr, per in getattr(m, tab).items():
                for sp, arr in per.items():
                    arrays["%%s/%%s/%%s/%%s" %% (name, tab, dr, sp)] = np.asarray(arr, dtype=np.float64)
    meta["models"][name] = mm
for k, p in comp.pin.items():
    meta["pins"][k] = {"model": p.model, "R": p.R, "L": p.L, "C": p.C}
meta["package"] = {"r_pkg": comp.package.r_pkg, "l_pkg": comp.package.l_pkg,
                   "c_pkg": comp.package.c_pkg}
np.savez_compressed(%r, meta=json.dumps(meta), **arrays)
print("ibis tables:", len(arrays), "arrays")
''' % (mod, os.path.join(HERE, "ibis_tables.npz"))
    subprocess.run([sys.executable, "-c", code], check=True, cwd=TMP)


def dump_waveforms():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import backends
    import circuits
    from eispice_b200 import units
    out = {}
    meta = {}
    for name, (build, an, _checks) in circuits.all_cases().items():
        nl = build().netlist()
        sim = backends.RefSim(nl)
        if an[0] == "op":
            data, var = sim.op()
        else:
            data, var = sim.tran(*[units.float(x) for x in an[1:]])
        keep = circuits.GOLDEN_COLUMNS.get(name)
        if keep is not None:
            idx = [0] + [var.index(k) for k in keep]
            data, var = data[:, idx], [var[i] for i in idx]
        out[name] = data
        meta[name] = {"variables": var, "analysis": list(an),
                      "all_columns": keep is None}
        print("%-18s %s" % (name, data.shape))
    np.savez_compressed(os.path.join(HERE, "ref_waveforms.npz"),
                        meta=json.dumps(meta), **out)


if __name__ == "__main__":
    warnings.filterwarnings("ignore")
    if not os.path.isdir(REF):
        sys.exit("make_golden: %s not present (fixtures are generated in the "
                 "build container only)" % REF)
    if "--waveforms-only" not in sys.argv:
        dump_ibis_tables(build_reference_python())
    dump_waveforms()
