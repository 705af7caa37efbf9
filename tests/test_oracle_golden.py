"""Pins the CPU oracle (oracle/eis_oracle.c) to the reference.

1. Against the committed golden waveforms produced by the UNMODIFIED reference
   engine (tests/golden/ref_waveforms.npz, made by tests/golden/make_golden.py):
   same number of accepted steps, same time points, values within the tight
   gate 1e-6*|ref| + 1e-9 for (nearly) every sample.
2. Against the reference's own known-answer doctest values (+/-0.01 %).
3. When oracle/_ref was built (build container), directly against the live
   reference engine as well.
"""
import json
import os

import numpy as np
import pytest

import backends
import circuits
import parity
from eispice_b200 import units

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_waveforms.npz"))
META = json.loads(str(GOLD["meta"]))

# Circuits whose step control is rounding-noise dominated in the reference
# itself (hundreds of ~1e-18 s steps at start-up / a 1-ulp final step): any
# change of LU rounding moves the accepted-step sequence.  They are held to the
# reference's own SPICE tolerance after interpolation, not to the tight gate.
SEQUENCE_SENSITIVE = {"ibis_ramp", "cfg2_rectifier"}
# minimum fraction of samples inside the tight gate (isolated ill-conditioned
# steps, e.g. h = 1e-23 s, amplify pivot-order rounding)
MIN_TIGHT_FRACTION = 0.995


def _run_oracle(name):
    build, an, checks = circuits.all_cases()[name]
    sim = backends.OracleSim(build().netlist())
    if an[0] == "op":
        data, var = sim.op()
    else:
        data, var = sim.tran(*[units.float(x) for x in an[1:]])
    return data, var, sim, an, checks


@pytest.mark.parametrize("name", sorted(circuits.all_cases()))
def test_oracle_matches_golden(name):
    data, var, sim, an, checks = _run_oracle(name)
    ref = GOLD[name]
    rvar = META[name]["variables"]
    idx = [var.index(v) for v in rvar]
    mine = data[:, idx]
    if name in SEQUENCE_SENSITIVE:
        tstep, tstop = units.float(an[1]), units.float(an[2])
        worst = parity.grid_error(mine[:, 0], mine, ref[:, 0], ref, rvar, tstep, tstop)
        assert worst < 10.0, "grid error %.3g x SPICE tolerance" % worst
        return
    assert mine.shape == ref.shape, "accepted-step count differs from the reference"
    np.testing.assert_allclose(mine[:, 0], ref[:, 0], rtol=1e-5, atol=0)  # LTE steps carry rounding noise
    frac, worst = parity.tight_fraction(mine, ref)
    assert frac >= MIN_TIGHT_FRACTION, "only %.4f of samples inside 1e-6/1e-9 (worst %.3g)" % (frac, worst)


@pytest.mark.parametrize("name", sorted(n for n, c in circuits.all_cases().items() if c[2]))
def test_oracle_known_answers(name):
    data, var, sim, an, checks = _run_oracle(name)
    for kind, node, value, time in checks:
        col = var.index("%s(%s)" % (kind, node))
        if len(data) > 1:
            got = float(np.interp(units.float(time), data[:, 0], data[:, col]))
        else:
            got = float(data[0, col])
        assert parity.reference_check(units.float(value), got), (kind, node, value, got)


@pytest.mark.skipif(not backends.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["cfg1_rc", "tline", "diode", "cfg3_oscillator", "cfg4_ibis"])
def test_oracle_matches_live_reference(name):
    build, an, _ = circuits.all_cases()[name]
    nl = build().netlist()
    a = [units.float(x) for x in an[1:]]
    do, vo = backends.OracleSim(nl).tran(*a)
    dr, vr = backends.RefSim(nl).tran(*a)
    assert vo == vr
    assert do.shape == dr.shape
    frac, worst = parity.tight_fraction(do, dr)
    assert frac >= MIN_TIGHT_FRACTION, (frac, worst)


@pytest.mark.skipif(not backends.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", sorted(circuits.EXTRA_CASES))
def test_oracle_matches_live_reference_extra(name):
    build, an, _ = circuits.EXTRA_CASES[name]
    nl = build().netlist()
    a = [units.float(x) for x in an[1:]]
    do, vo = backends.OracleSim(nl).tran(*a)
    dr, vr = backends.RefSim(nl).tran(*a)
    assert vo == vr
    assert do.shape == dr.shape
    frac, worst = parity.tight_fraction(do, dr)
    assert frac >= MIN_TIGHT_FRACTION, (frac, worst)


def test_oracle_step_statistics():
    """k = solves per accepted step, as probed on the reference."""
    _, _, sim, _, _ = _run_oracle("cfg4_ibis")
    st = sim.stats()
    assert st["accepted"] == 3422 and st["rejected_lte"] == 913
    assert abs(st["factorizations"] / st["accepted"] - 5.0) < 0.05


def test_parameter_sweep_matches_rebuild():
    """eoSetParam + restart equals building the circuit with that value."""
    a = backends.OracleSim(circuits.cfg1_rc().netlist())
    a.tran(0.01e-9, 10e-9)
    a.set_param("R1", "R", 2.5e3)
    d1, _ = a.tran(0.01e-9, 10e-9, restart=True)
    cct = circuits.cfg1_rc()
    cct.R1.R = 2.5e3
    d2, _ = backends.OracleSim(cct.netlist()).tran(0.01e-9, 10e-9)
    np.testing.assert_array_equal(d1, d2)


@pytest.mark.skipif(not backends.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("cfg,index", [("cfg2", 17), ("cfg3", 5), ("cfg4", 4), ("cfg4", 9), ("cfg5", 2)])
def test_oracle_matches_live_reference_on_sweep_instances(cfg, index):
    """The oracle is pinned on the SWEEP space too: instances of the synthetic
    workloads (random parameters, non-typical IBIS corners, falling edges) run on
    the unmodified reference and on the oracle."""
    import workloads
    w = workloads.make(cfg, 24)
    nl, s = w.instance_netlist(index)
    do, vo = backends.OracleSim(nl, s).tran(w.tstep, w.tstop)
    dr, vr = backends.RefSim(nl, s).tran(w.tstep, w.tstop)
    assert vo == vr
    do, dr = parity.drop_degenerate(do, w.tstep), parity.drop_degenerate(dr, w.tstep)
    worst = parity.grid_error(do[:, 0], do, dr[:, 0], dr, vo, w.tstep, w.tstop, current_floor=1e-3)
    assert worst < 5.0, "grid error %.3g x SPICE tolerance" % worst
    if do.shape == dr.shape and cfg != "cfg2":     # cfg2 is SEQUENCE_SENSITIVE (see top)
        frac, tight = parity.tight_fraction(do, dr)
        assert frac >= 0.98, (frac, tight)
