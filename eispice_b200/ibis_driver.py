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
