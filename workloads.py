"""Synthetic sweeps of the five BASELINE.json configurations,
shared by bench.py, the GPU tests and __graft_entry__.smoke().

A workload is a topology (an ``eispice_b200.Circuit``), per-instance parameter
arrays keyed "Refdes.member" (plus "Refdes.set" = IBIS table-set index), the
probes and the analysis window.  Seeds: numpy default_rng(20260 + cfg).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import circuits  # noqa: E402  (tests/circuits.py: the golden circuit builders)
from eispice_b200 import ibis_device  # noqa: E402

DEFAULT_INSTANCES = {"cfg1": 65536, "cfg2": 10000, "cfg3": 100000, "cfg4": 262144,
                     "cfg5": 131072}
DESCRIPTION = {
    "cfg1": "RC low-pass: pulse V + R + C, tran 0.01n/10n",
    "cfg2": "half-wave diode rectifier with RC load, Rl x Cl grid sweep, tran 10n/3u",
    "cfg3": "B-source nonlinear LC oscillator, L/C Monte-Carlo, tran 1n/1u",
    "cfg4": "IBIS driver + lossless T-line + IBIS receiver, corner x Rt/Z0/Td sweep, tran 0.01n/15n",
    "cfg5": "IBIS driver + 8 lossy T-line sections + receiver Monte-Carlo, tran 0.01n/15n",
}


class Workload:
    def __init__(self, name, cct, params, probes, tstep, tstop, sets=None):
        self.name = name
        self.cct = cct
        self.params = params          # {"Refdes.member": array[M]}
        self.probes = probes
        self.tstep = tstep
        self.tstop = tstop
        self.sets = sets              # [(speed, direction)] when table sets are swept
        self.M = len(next(iter(params.values())))

    def instance_netlist(self, i):
        """Flat netlist of instance i alone (for the CPU oracle / reference)."""
        nl = self.cct.netlist()
        scal = {}
        vi_set = 0
        for key, vals in self.params.items():
            ref, member = key.rsplit(".", 1)
            if member == "set":
                vi_set = int(vals[i])
            else:
                scal[(ref, member)] = float(vals[i])
        out = []
        for d in nl:
            d = list(d)
            k, ref = d[0], d[1]
            if k in ("R", "C", "L") and (ref, k) in scal:
                d[4] = scal[(ref, k)]
            elif k == "T":
                for j, m in enumerate(("Z0", "Td", "loss")):
                    if (ref, m) in scal:
                        d[6 + j] = scal[(ref, m)]
            elif k in ("V", "I") and (ref, "DC") in scal:
                d[5] = scal[(ref, "DC")]
            out.append(tuple(d))
        return out, vi_set


def _grid(M):
    side = int(np.ceil(np.sqrt(M)))
    a, b = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    return side, a.ravel()[:M], b.ravel()[:M]


def cfg1(M, seed=20261):
    rng = np.random.default_rng(seed)
    cct = circuits.cfg1_rc()
    params = {"R1.R": 1e3 * rng.uniform(0.5, 2.0, M), "C1.C": 1e-12 * rng.uniform(0.5, 2.0, M)}
    if M == 1:
        params = {"R1.R": np.array([1e3]), "C1.C": np.array([1e-12])}
    return Workload("cfg1", cct, params, ["v(out)", "i(V1)"], 0.01e-9, 10e-9)


def cfg2(M, seed=20262):
    cct = circuits.cfg2_rectifier()
    side, a, b = _grid(M)
    rl = np.logspace(2, 4, side)[a]
    cl = np.logspace(-7, -5, side)[b]
    return Workload("cfg2", cct, {"Rl.R": rl, "Cl.C": cl}, ["v(out)", "i(V1)"], 10e-9, 3e-6)


def cfg3(M, seed=20263):
    rng = np.random.default_rng(seed)
    cct = circuits.cfg3_oscillator()
    l1 = 1e-6 * (1 + 0.05 * np.clip(rng.standard_normal(M), -3, 3))
    c1 = 1e-9 * (1 + 0.05 * np.clip(rng.standard_normal(M), -3, 3))
    return Workload("cfg3", cct, {"L1.L": l1, "C1.C": c1}, ["v(a)", "i(L1)"], 1e-9, 1e-6)


_CORNERS = ibis_device.corner_sets(["typ", "min", "max"], ["rising", "falling"])


def _ibis_link(sections):
    """cfg4/cfg5 topology with all six (speed, direction) table sets attached."""
    import eispice_b200 as eispice
    cct = circuits._new("IBIS link")
    models, pins = circuits.ibis_fixture()
    speeds, dirs = ["typ", "min", "max"], ["rising", "falling"]
    drv = pins['2']
    rcv = pins['1']
    cct.Driver = ibis_device.Pin(models[drv["model"]], drv, 'vsx', speeds, dirs)
    cct.Vmeas = eispice.V('vsx', 'vs', 0)
    cct.Rt = eispice.R('vs', 'vi', 33.2)
    if sections is None:
        cct.Tg = eispice.T('vi', 0, 'vo', 0, 50, '2n')
    else:
        prev = 'vi'
        for k in range(sections):
            nxt = 'vo' if k == sections - 1 else 'n%d' % (k + 1)
            setattr(cct, 'T%d' % (k + 1), eispice.T(prev, 0, nxt, 0, 50, 0.25e-9, 0.05))
            prev = nxt
    cct.Receiver = ibis_device.Pin(models[rcv["model"]], rcv, 'vo', speeds, dirs)
    return cct


def _set_params(cct, sets):
    """Per-instance corner: "Refdes.set" for every VI device carrying table
    sets, and the speed-dependent scalars of the pins (supply voltage, C_comp:
    reference module/ibis_device.py:103,218)."""
    models, pins = circuits.ibis_fixture()
    speed_of = [s for s, _ in _CORNERS]
    out = {}
    pin_models = []          # in subckt-id order: driver pin '2', receiver pin '1'
    for d in cct.netlist():
        if d[0] == "VI" and len(d[4]) > 1:
            out["%s.set" % d[1]] = sets
        if d[0] == "V" and d[1].endswith("#Vcc"):
            model = models[pins['2' if not pin_models else '1']["model"]]
            pin_models.append(model)
            vr = np.array([model["voltage_range"][s] for s in speed_of])
            out["%s.DC" % d[1]] = vr[sets]
        if d[0] == "C" and d[1].endswith("#C_comp"):
            cc = np.array([pin_models[-1]["c_comp"][s] for s in speed_of])
            out["%s.C" % d[1]] = cc[sets]
    return out


def cfg4(M, seed=20264):
    rng = np.random.default_rng(seed)
    cct = _ibis_link(None)
    sets = (np.arange(M) % len(_CORNERS)).astype(np.int32)
    params = _set_params(cct, sets)
    params.update({"Rt.R": rng.uniform(20, 50, M), "Tg.Z0": rng.uniform(40, 60, M),
                   "Tg.Td": rng.uniform(1.5e-9, 2.5e-9, M)})
    return Workload("cfg4", cct, params, ["v(vo)", "i(Vmeas)"], 0.01e-9, 15e-9, _CORNERS)


def cfg5(M, seed=20265):
    rng = np.random.default_rng(seed)
    cct = _ibis_link(8)
    sets = (np.arange(M) % len(_CORNERS)).astype(np.int32)
    params = _set_params(cct, sets)
    for k in range(1, 9):
        params["T%d.Z0" % k] = 50 + 2.5 * np.clip(rng.standard_normal(M), -3, 3)
        params["T%d.Td" % k] = 0.25e-9 * (1 + 0.05 * rng.uniform(-1, 1, M))
        params["T%d.loss" % k] = 0.05 * (1 + 0.2 * rng.uniform(-1, 1, M))
    return Workload("cfg5", cct, params, ["v(vo)", "i(Vmeas)"], 0.01e-9, 15e-9, _CORNERS)


BUILDERS = {"cfg1": cfg1, "cfg2": cfg2, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5}


def make(name, M=None, seed_offset=0):
    M = DEFAULT_INSTANCES[name] if M is None else int(M)
    fn = BUILDERS[name]
    base = {"cfg1": 20261, "cfg2": 20262, "cfg3": 20263, "cfg4": 20264, "cfg5": 20265}[name]
    return fn(M, base + seed_offset)


def algorithmic_bytes(stats, n_tlines, n_outputs):
    """B_alg of SURVEY.md 8d per accepted instance-timestep, from the engine's
    own counters: 8*[k(nnzA + 2F + 2N) + 2(nnzA + N) + 2S + H + O]."""
    k = stats["factorizations"] / max(1, stats["accepted"])
    N, nnz, F, S = (stats["numUnknowns"], stats["numNonzeros"], stats["numFactor"],
                    stats["stateDoubles"])
    H = 18 * n_tlines
    return 8.0 * (k * (nnz + 2 * F + 2 * N) + 2 * (nnz + N) + 2 * S + H + n_outputs), k


def cfg5_masked(M, seed=20266, max_sections=8):
    """the cfg5 link as a MAXIMUM topology of 8 sections, of
    which instance i uses only n_i = 1 + i % 8 - the others are bypassed
    ("Tk.bypass" = 1: an ideal through-connection).  One topology, one kernel."""
    w = cfg5(M, seed)
    n_active = 1 + (np.arange(M) % max_sections)
    for k in range(1, max_sections + 1):
        w.params["T%d.bypass" % k] = (k > n_active).astype(np.float64)
    w.name = "cfg5_masked"
    w.n_active = n_active
    return w


def reduced_link_netlist(w, i):
    """Instance i of cfg5_masked as the REDUCED topology the reference can run:
    the link built with only its n_i active sections, same parameters."""
    n = int(w.n_active[i])
    cct = _ibis_link(n)
    full = {k: v[i] for k, v in w.params.items()}
    vi_set = int(full[[k for k in full if k.endswith(".set")][0]])
    out = []
    for d in cct.netlist():
        d = list(d)
        k, ref = d[0], d[1]
        if k in ("R", "C", "L") and "%s.%s" % (ref, k) in full:
            d[4] = float(full["%s.%s" % (ref, k)])
        elif k == "T":
            for j, m in enumerate(("Z0", "Td", "loss")):
                d[6 + j] = float(full["%s.%s" % (ref, m)])
        elif k in ("V", "I") and "%s.DC" % ref in full:
            d[5] = float(full["%s.DC" % ref])
        out.append(tuple(d))
    return out, vi_set
