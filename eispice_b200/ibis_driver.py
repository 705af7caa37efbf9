"""IBIS switching-coefficient (K) tables, mirroring the reference's
``module/ibis_driver.py``: the time/multiplier tables ku(t), kd(t) that turn a
driver's static pull-up / pull-down V-I tables into a switching buffer are
computed by running a small transient of the buffer against its published
V-t waveforms (``calcDriverKDoubleWave`` :40-203) or, when fewer than two
waveforms exist, against a Gaussian edge built from the ramp data
(``calcDriverKRamp`` :321-435; ``calcDriverKSingleWave`` :205-319 falls back to
it).  The circuits are built exactly as the reference builds them; the
transient runs on the batched CUDA engine.

A model is a plain dictionary (the layout of eispice_b200.ibis_device) with

  "rising_waveform" / "falling_waveform":
        [{"data": {speed: [[t, v], ...]}, "r_fixture": ohm, "v_fixture": V,
          ("v_fixture_min": V, "v_fixture_max": V)}, ...]
  "ramp": {"dt_r": {speed: s}, "dt_f": {speed: s}, "r_load": ohm}
"""
import numpy as np

from . import circuit, device, waveform

Typical, Minimum, Maximum = "typ", "min", "max"
Rising, Falling = "rising", "falling"


def _v_fixture(wave, speed):
    """reference ibis_driver.py:126-137: max speed pairs with v_fixture_min."""
    if speed == Maximum:
        return wave.get("v_fixture_min", wave["v_fixture"])
    if speed == Minimum:
        return wave.get("v_fixture_max", wave["v_fixture"])
    return wave["v_fixture"]


def _half(cct, model, speed, wave, tag):
    """One instantiation of the buffer against one V-t waveform (:117-143)."""
    wv, ts, fix = "wv" + tag, "ts" + tag, "fix" + tag
    setattr(cct, "Xwv" + tag, device.V(wv, 0, 0, waveform.PWL(wave["data"][speed])))
    setattr(cct, "Xpu" + tag, device.VI("vcc", wv, waveform.PWL(model["pullup"][speed])))
    setattr(cct, "Xpd" + tag, device.VI(wv, 0, waveform.PWL(model["pulldown"][speed])))
    setattr(cct, "Vmeas" + tag, device.V(wv, ts, 0))
    setattr(cct, "Cc" + tag, device.C(ts, 0, model["c_comp"][speed]))
    setattr(cct, "Rfix" + tag, device.R(ts, fix, wave["r_fixture"]))
    setattr(cct, "Vfix" + tag, device.V(fix, 0, _v_fixture(wave, speed)))
    if "power_clamp" in model:
        setattr(cct, "Xcu" + tag, device.VI("vcc", ts, waveform.PWL(model["power_clamp"][speed])))
    if "gnd_clamp" in model:
        setattr(cct, "Xcd" + tag, device.VI(ts, 0, waveform.PWL(model["gnd_clamp"][speed])))


def double_wave_circuit(model, direction, speed, tstep_divisor=10):
    """(circuit, tstep, tstop) of calcDriverKDoubleWave (ibis_driver.py:40-203)."""
    wave = model[direction + "_waveform"]
    cct = circuit.Circuit("Double Waveform")
    cct.Vcc = device.V("vcc", 0, model["voltage_range"][speed])
    _half(cct, model, speed, wave[0], "0")
    _half(cct, model, speed, wave[1], "1")
    cct.Bkdn = device.B("kd", 0, "v",
                        "(i(Xpu0)*i(Vmeas1) - i(Xpu1)*i(Vmeas0))"
                        "/ (i(Xpd0)*i(Xpu1) - i(Xpd1)*i(Xpu0))")
    cct.Bkun = device.B("ku", 0, "v",
                        "(i(Xpd0)*i(Vmeas1) - i(Xpd1)*i(Vmeas0))"
                        "/ (i(Xpd0)*i(Xpu1) - i(Xpd1)*i(Xpu0))")
    d0 = np.asarray(wave[0]["data"][speed])
    d1 = np.asarray(wave[1]["data"][speed])
    tstop = max(d0[-1][0], d1[-1][0])
    tstep = min(d0[1][0] - d0[0][0], d1[1][0] - d1[0][0])
    return cct, tstep / tstep_divisor, tstop


def ramp_circuit(model, direction, speed, tstep_divisor=10):
    """(circuit, tstep, tstop) of calcDriverKRamp (ibis_driver.py:321-435)."""
    vol, voh = 0.0, model["voltage_range"][speed]
    if direction == Rising:
        ramp = model["ramp"]["dt_r"][speed]
        wave = waveform.Gauss(vol, voh, 0, ramp)
    else:
        ramp = model["ramp"]["dt_f"][speed]
        wave = waveform.Gauss(voh, vol, 0, ramp)
    cct = circuit.Circuit("Ramp")
    cct.Vcc = device.V("vcc", 0, model["voltage_range"][speed])
    cct.Xwv = device.V("wv", 0, 0, wave)
    cct.Xpu = device.VI("vcc", "wv", waveform.PWL(model["pullup"][speed]))
    cct.Xpd = device.VI("wv", 0, waveform.PWL(model["pulldown"][speed]))
    cct.Vmeas = device.V("wv", "ts", 0)
    cct.Cc = device.C("ts", 0, model["c_comp"][speed])
    cct.Rfix = device.R("ts", "fix", model["ramp"]["r_load"])
    cct.Vfix = device.V("fix", 0, 0.0 if direction == Rising else model["voltage_range"][speed])
    if "power_clamp" in model:
        cct.Xcu = device.VI("vcc", "ts", waveform.PWL(model["power_clamp"][speed]))
    if "gnd_clamp" in model:
        cct.Xcd = device.VI("ts", 0, waveform.PWL(model["gnd_clamp"][speed]))
    cct.Bkdn = device.B("kd", 0, "v", "(i(Vmeas) - i(Xpu) - i(Vfix))/ (i(Xpd) - i(Xpu))")
    cct.Bkun = device.B("ku", 0, "v", "(i(Vmeas) - i(Xpd) - i(Vfix))/ (i(Xpd) - i(Xpu))")
    return cct, ramp * 0.5 / tstep_divisor, ramp * 4


def k_circuit(model, direction, speed, tstep_divisor=10):
    """The circuit the reference picks for this model (ibis_parser.py:481-497):
    two waveforms -> double-wave method, otherwise the ramp."""
    waves = model.get(direction + "_waveform") or []
    if len(waves) >= 2:
        return double_wave_circuit(model, direction, speed, tstep_divisor)
    return ramp_circuit(model, direction, speed, tstep_divisor)


def _gpu_runner(cct, tstep, tstop):
    cct.tran(tstep, tstop)
    return cct.results, cct.variables


def calc_driver_k(model, direction, speed, tstep_divisor=10, runner=None):
    """(ku, kd): [[t, k], ...] arrays without the first point, as the reference
    returns them (ibis_driver.py:196-201).  ``runner(cct, tstep, tstop)`` ->
    (results, variables) defaults to the CUDA engine."""
    cct, tstep, tstop = k_circuit(model, direction, speed, tstep_divisor)
    data, names = (runner or _gpu_runner)(cct, tstep, tstop)
    t = names.index("time")
    ku = data[:, [t, names.index("v(ku)")]][1:]
    kd = data[:, [t, names.index("v(kd)")]][1:]
    return ku, kd


def driver_k_tables(model, runner=None):
    """pullup_k / pulldown_k of a driver model for both directions and the three
    speeds (ibis_parser.py:500-510), in the layout ibis_device expects."""
    up, down = {}, {}
    for direction in (Rising, Falling):
        up[direction], down[direction] = {}, {}
        for speed in (Typical, Minimum, Maximum):
            ku, kd = calc_driver_k(model, direction, speed, runner=runner)
            up[direction][speed] = ku
            down[direction][speed] = kd
    return {"pullup_k": up, "pulldown_k": down}


# ---------------------------------------------------------------------------
# Batched generation: every (model, direction, speed) transient is one INSTANCE
# of a batch per extraction-circuit topology (reference: 6 separate transients
# per driver model at parse time, module/ibis_parser.py:500-510).
# ---------------------------------------------------------------------------
def _signature(netlist):
    """What must be equal for two extraction circuits to share a batch: device
    kinds, names, nodes, B equations, waveform kinds and which VI devices carry a
    TA table - everything except values and table contents."""
    sig = []
    for d in netlist:
        k = d[0]
        if k in ("R", "C", "L"):
            sig.append((k, d[1], d[2], d[3]))
        elif k in ("V", "I"):
            sig.append((k, d[1], d[2], d[3], None if d[6] is None else d[6][0]))
        elif k == "B":
            sig.append((k, d[1], d[2], d[3], d[4], d[5]))
        elif k == "VI":
            sig.append((k, d[1], d[2], d[3], tuple(t[0] for t in d[4]), d[5] is not None))
        else:
            sig.append(tuple(d[:6]))
    return tuple(sig)


def _table_key(t):
    return np.ascontiguousarray(np.asarray(t, dtype=np.float64)).tobytes()


def merge_netlists(netlists):
    """Several netlists of one signature -> (netlist with table sets, params):
    scalars that differ become per-instance arrays "Refdes.member", tables that
    differ become table sets selected by "Refdes.set"."""
    base = netlists[0]
    n = len(netlists)
    merged, params = [], {}
    for k, d0 in enumerate(base):
        devs = [nl[k] for nl in netlists]
        kind, ref = d0[0], d0[1]
        if kind in ("R", "C", "L"):
            vals = np.array([float(d[4]) for d in devs])
            if (vals != vals[0]).any():
                params["%s.%s" % (ref, kind)] = vals
            merged.append(d0)
        elif kind in ("V", "I"):
            dc = np.array([float(d[5]) for d in devs])
            if (dc != dc[0]).any():
                params["%s.DC" % ref] = dc
            wave = d0[6]
            if wave is None:
                merged.append(d0)
            elif wave[0] in ("l", "c"):
                keys, tables, sel = {}, [], []
                for d in devs:
                    key = _table_key(d[6][1])
                    if key not in keys:
                        keys[key] = len(tables)
                        tables.append(d[6][1])
                    sel.append(keys[key])
                if len(tables) > 1:
                    params["%s.set" % ref] = np.array(sel, dtype=np.int32)
                merged.append(tuple(d0[:6]) + ((wave[0], tables[0], tables[1:]),))
            else:
                args = np.array([[float(x) for x in d[6][1]] for d in devs])
                for j in range(args.shape[1]):
                    col = args[:, j]
                    same = (col == col[0]) | (np.isinf(col) & np.isinf(col[0]))
                    if not same.all():
                        params["%s.A%d" % (ref, j)] = col
                merged.append(d0)
        elif kind == "VI":
            vi_keys, vis, tas, sel = {}, [], [], []
            for d in devs:
                key = _table_key(d[4][0][1]) + (b"" if d[5] is None else _table_key(d[5][0][1]))
                if key not in vi_keys:
                    vi_keys[key] = len(vis)
                    vis.append(d[4][0])
                    if d[5] is not None:
                        tas.append(d[5][0])
                sel.append(vi_keys[key])
            if len(vis) > 1:
                params["%s.set" % ref] = np.array(sel, dtype=np.int32)
            merged.append((kind, ref, d0[2], d0[3], vis, tas if d0[5] is not None else None))
        else:
            merged.append(d0)
    for key in params:
        assert len(params[key]) == n
    return merged, params


def instance_netlist(merged, params, j):
    """Instance j of a merged netlist as a plain netlist (checks merge_netlists;
    also what the CPU oracle would run for that instance)."""
    out = []
    for d in merged:
        d = list(d)
        kind, ref = d[0], d[1]
        if kind in ("R", "C", "L") and "%s.%s" % (ref, kind) in params:
            d[4] = float(params["%s.%s" % (ref, kind)][j])
        elif kind in ("V", "I"):
            if "%s.DC" % ref in params:
                d[5] = float(params["%s.DC" % ref][j])
            wave = d[6]
            if wave is not None and wave[0] in ("l", "c"):
                tables = [wave[1]] + list(wave[2] if len(wave) > 2 else [])
                s = int(params["%s.set" % ref][j]) if "%s.set" % ref in params else 0
                d[6] = (wave[0], tables[s])
            elif wave is not None:
                args = list(wave[1])
                for a in range(len(args)):
                    if "%s.A%d" % (ref, a) in params:
                        args[a] = float(params["%s.A%d" % (ref, a)][j])
                d[6] = (wave[0], args)
        elif kind == "VI":
            s = int(params["%s.set" % ref][j]) if "%s.set" % ref in params else 0
            d[4] = [d[4][s]]
            d[5] = None if d[5] is None else [d[5][s]]
        out.append(tuple(d))
    return out


def driver_k_tables_batched(models, tstep_divisor=10, max_points=60000):
    """pullup_k / pulldown_k of every driver model in ``models`` ({name: model}):
    all their (direction, speed) transients run as instances of ONE batch per
    extraction-circuit topology on the CUDA engine, each with its own analysis
    window (simcuSetInstanceWindow), its own V-t waveform / V-I tables (table
    sets) and its own scalars.  Returns ({name: {"pullup_k":…, "pulldown_k":…}},
    info) with info = number of batches, instances and accepted steps."""
    from . import engine
    jobs = []
    for name, model in models.items():
        for direction in (Rising, Falling):
            for speed in (Typical, Minimum, Maximum):
                cct, tstep, tstop = k_circuit(model, direction, speed, tstep_divisor)
                jobs.append((name, direction, speed, cct.netlist(), tstep, tstop))
    groups = {}
    for job in jobs:
        groups.setdefault(_signature(job[3]), []).append(job)
    out = {name: {"pullup_k": {Rising: {}, Falling: {}}, "pulldown_k": {Rising: {}, Falling: {}}}
           for name in models}
    info = {"batches": 0, "instances": 0, "accepted": 0, "failed": 0}
    for group in groups.values():
        merged, params = merge_netlists([g[3] for g in group])
        n = len(group)
        eng = engine.Engine(merged, n)
        try:
            if params:
                eng.set_params(params)
            tsteps = np.array([g[4] for g in group])
            tstops = np.array([g[5] for g in group])
            eng.set_windows(tsteps, tstops)
            res = eng.tran_batch(float(np.median(tsteps)), float(tstops.max()), 0.0, None, [],
                                 raw_max_points=max_points, raw_count=n)
            raws = eng.fetch_raw_all(max_points)
            names = eng.variables()
        finally:
            eng.close()
        t, iu, idn = names.index("time"), names.index("v(ku)"), names.index("v(kd)")
        for j, (name, direction, speed, _, _, _) in enumerate(group):
            if res.status[j] != 0:
                info["failed"] += 1
                continue
            out[name]["pullup_k"][direction][speed] = raws[j][:, [t, iu]][1:]
            out[name]["pulldown_k"][direction][speed] = raws[j][:, [t, idn]][1:]
        info["batches"] += 1
        info["instances"] += n
        info["accepted"] += int(res.stats["accepted"])
    return out, info
