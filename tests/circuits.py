"""Golden circuits: every transient / operating-point doctest of the reference
that exercises the hot path, plus the five BASELINE.json
configurations.  Each builder returns an ``eispice_b200.Circuit``; ``CASES``
adds the analysis and the reference's own known-answer checks
(kind, name, value, time) copied from the cited docstrings.
"""
import numpy as np

import eispice_b200 as eispice
from eispice_b200 import subckt


def _new(title):
    subckt.reset_counter()
    return eispice.Circuit(title)


def inductor():                      # reference module/device.py:72-79
    cct = _new("Inductor Test")
    wave = eispice.Pulse(4, 8, '10n', '2n', '3n', '5n', '20n')
    cct.Ix = eispice.I(1, eispice.GND, 4, wave)
    cct.Lx = eispice.L(1, eispice.GND, '10n')
    return cct


def capacitor():                     # device.py:96-104
    cct = _new("Capacitor Test")
    wave = eispice.Pulse(4, 8, '10n', '2n', '3n', '5n', '20n')
    cct.Vx = eispice.V(1, eispice.GND, 4, wave)
    cct.Cx = eispice.C(1, eispice.GND, '10n')
    return cct


def resistor():                      # device.py:120-128
    cct = _new("Resistor Test")
    cct.Rv = eispice.R(1, eispice.GND, 324)
    cct.Vx = eispice.V(1, eispice.GND, 12.6)
    return cct


def realc():                         # device.py:146-153
    cct = _new("Real Capacitor Test")
    wave = eispice.Pulse(4, 8, '10n', '2n', '3n', '5n', '20n')
    cct.Vx = eispice.V(1, eispice.GND, 4, wave)
    cct.Cx = eispice.RealC(1, eispice.GND, '10n', '1n', '0.1')
    return cct


def b_current():                     # device.py:203-217
    cct = _new("Nonlinear Current Test")
    cct.Rv = eispice.R(1, eispice.GND, 10)
    cct.Vx = eispice.V(1, eispice.GND, 7)
    cct.Rb = eispice.R(2, eispice.GND, 3)
    cct.Bx = eispice.B(2, eispice.GND, eispice.Current,
                       "4.739057e-04 * (uramp( v(2,0) --5.060000e+00)) "
                       "+ sqrt(v(1)) / sinh(i(Vx))")
    return cct


def b_voltage_time():                # device.py:220-229
    cct = _new("Nonlinear Time Test")
    cct.Rb = eispice.R(2, eispice.GND, 3)
    cct.Bx = eispice.B(2, eispice.GND, eispice.Voltage,
                       'sin(2*3.14159*100e6*time)')
    return cct


def b_capacitor():                   # device.py:232-243
    cct = _new("Non-Linear Capacitor Test")
    wave = eispice.Pulse('1u', '10u', '10n', '5n', '3n', '5n', '50n')
    cct.Vc = eispice.V(3, 0, '1n', wave)
    cct.Vx = eispice.V(1, 0, 1,
                       eispice.Pulse('1m', '10m', '10n', '2n', '3n', '5n', '20n'))
    cct.Cx = eispice.B(1, 0, eispice.Capacitor, 'v(3)*10')
    return cct


def i_pulse():                       # device.py:401-410
    cct = _new("Current Pulse Test")
    cct.Ix = eispice.I(1, eispice.GND, 4,
                       eispice.Pulse(4, 8, '10n', '2n', '3n', '5n', '20n'))
    cct.Vx = eispice.V(2, 1, 0)
    cct.Rx = eispice.R(2, eispice.GND, 10)
    return cct


def v_pulse():                       # device.py:432-440
    cct = _new("Voltage Pulse Test")
    cct.Ix = eispice.V(1, eispice.GND, 4,
                       eispice.Pulse(4, 8, '10n', '2n', '3n', '5n', '20n'))
    cct.Rx = eispice.R(1, eispice.GND, 10)
    return cct


def vi_table(dc=10):                 # device.py:462-473
    cct = _new("VI Curve Test")
    data = np.array([[-10, -10], [-5, -2], [0, 1], [1, 3], [5, 8],
                     [10, 10], [12, 8]], dtype=float)
    cct.Vx = eispice.V(1, 0, dc)
    cct.VIx = eispice.VI(1, 0, eispice.PWL(data))
    return cct


def vccs():                          # device.py:494-501
    cct = _new("VCCS Test")
    cct.Vx = eispice.V(1, 0, 3.2)
    cct.Iy = eispice.G(2, 0, 1, 0, 2)
    cct.Ry = eispice.R(2, 3, 4.1)
    cct.Vy = eispice.V(3, 0, 0)
    return cct


def vcvs():                          # device.py:522-528
    cct = _new("VCVS Test")
    cct.Vx = eispice.V(1, 0, 3.2)
    cct.Vy = eispice.E(2, 0, 1, 0, 2.7)
    cct.Ry = eispice.R(2, 0, 4.1)
    return cct


def cccs():                          # device.py:549-558
    cct = _new("CCCS Test")
    cct.Ix = eispice.I(1, 0, 3.2)
    cct.Rx = eispice.R(1, 4, 7.3)
    cct.Vx = eispice.V(4, 0, 0)
    cct.Iy = eispice.F(2, 0, 'Vx', 1.75)
    cct.Ry = eispice.R(2, 3, 6.2)
    cct.Vy = eispice.V(3, 0, 0)
    return cct


def ccvs():                          # device.py:578-586
    cct = _new("CCVS Test")
    cct.Ix = eispice.I(1, 0, 2.7)
    cct.Rx = eispice.R(1, 4, 10.4)
    cct.Vx = eispice.V(4, 0, 0)
    cct.Vy = eispice.H(2, 0, 'Vx', 3.2)
    cct.Ry = eispice.R(2, 0, 3.7)
    return cct


def tline():                         # device.py:610-620
    cct = _new("Simple Transmission Line Model Test")
    cct.Vx = eispice.V('vs', 0, 0, eispice.Pulse(0, 1, '0n', '1n', '1n', '4n', '8n'))
    cct.Rt = eispice.R('vs', 'vi', 50)
    cct.Tg = eispice.T('vi', 0, 'vo', 0, 50, '2n')
    cct.Cx = eispice.C('vo', 0, '5p')
    return cct


def tline_lossy():                   # libs/simulator/tester.c:130-168 shape
    cct = _new("Lossy Transmission Line")
    cct.Vx = eispice.V('vs', 0, 0, eispice.Pulse(0, 1, '0n', '1n', '1n', '4n', '8n'))
    cct.Rt = eispice.R('vs', 'vi', 50)
    cct.Tg = eispice.T('vi', 0, 'vo', 0, 50, '2n', 0.2)
    cct.Rl = eispice.R('vo', 0, 75)
    return cct


def diode():                         # device.py:754-763
    cct = _new("Diode Test")
    wave = eispice.Pulse(0, 1, '.5u', '5u', '5u', '.5u')
    cct.Vx = eispice.V(1, eispice.GND, 0, wave)
    cct.Dx = eispice.D(1, eispice.GND, IS='2.52n', RS=0.568, N=1.752,
                       CJO='4p', M=0.4, TT='20n')
    return cct


def _wave_circuit(title, wave):      # waveform.py doctests share this shape
    cct = _new(title)
    cct.Rv = eispice.R(1, eispice.GND, 1)
    cct.Vx = eispice.V(1, 0, eispice.GND, wave)
    return cct


def w_pwl():                         # waveform.py:52-61
    return _wave_circuit('PWL Test', eispice.PWL(
        [['2n', 4], ['12n', 3], ['50n', 20], ['75n', -20], ['95n', -22]]))


def w_pwc():                         # waveform.py:80-89
    return _wave_circuit('PWC Test', eispice.PWC(
        [['2n', 4], ['12n', 3], ['50n', 20], ['75n', -20], ['95n', -22]]))


def w_sffm():                        # waveform.py:105-113
    return _wave_circuit('SFFM Test', eispice.SFFM(1, 4, '100M', 2, '10M'))


def w_exp():                         # waveform.py:131-139
    return _wave_circuit('Exp Test', eispice.Exp(0, 4, '5n', '2n', '25n', '5n'))


def w_pulse():                       # waveform.py:158-166
    return _wave_circuit('Pulse Test',
                         eispice.Pulse(4, 8, '10n', '2n', '3n', '5n', '20n'))


def w_gauss():                       # waveform.py:186-194
    return _wave_circuit('Gauss Test',
                         eispice.Gauss(0, 3.3, '0n', '2n', '5n', '10n', '50n'))


def w_sin():                         # waveform.py:214-222
    return _wave_circuit('Sin Test', eispice.Sin(0, 4, '50M', '5n', '10M'))


def subckt_nested():                 # subckt.py:38-51
    class Subckt1(eispice.Subckt):
        def __init__(self, pNode, nNode, Rx, Ry):
            self.Rx = eispice.R(pNode, self.node('node'), Rx)
            self.Ry = eispice.R(nNode, self.node('node'), Ry)

    class Subckt2(eispice.Subckt):
        def __init__(self, pNode, nNode, Rx, Ry):
            self.Rx = eispice.R(pNode, self.node('node'), Rx)
            self.Xy = Subckt1(nNode, self.node('node'), Rx, Ry)

    cct = _new("Subckt Test")
    cct.Vx = eispice.V(1, 0, 10)
    cct.Xx = Subckt2(1, 0, 100, 100)
    return cct


def circuit_basic():                 # circuit.py:106-117
    cct = _new("Circuit Test")
    cct.Vx = eispice.V(1, eispice.GND, 10)
    cct.Rx = eispice.R(1, eispice.GND, '100')
    return cct


def circuit_tran():                  # circuit.py:207-213
    cct = _new("Circuit Tran Test")
    cct.Vx = eispice.V(1, eispice.GND, 1)
    cct.Rx = eispice.R(1, eispice.GND, '1')
    return cct


# ---- BASELINE.json configurations ---------------------------
def cfg1_rc():
    cct = _new("cfg1 RC low-pass")
    cct.V1 = eispice.V('in', 0, 0, eispice.Pulse(0, 1, '1n', '0.1n', '0.1n', '4n', '10n'))
    cct.R1 = eispice.R('in', 'out', '1k')
    cct.C1 = eispice.C('out', 0, '1p')
    return cct


def cfg2_rectifier(Rl=1e3, Cl=1e-6):
    cct = _new("cfg2 half-wave rectifier")
    cct.V1 = eispice.V('in', 0, 0, eispice.Sin(0, 5, '1Meg'))
    cct.D1 = eispice.D('in', 'out', IS='2.52n', RS=0.568, N=1.752, CJO='4p',
                       M=0.4, TT='20n')
    cct.Rl = eispice.R('out', 0, Rl)
    cct.Cl = eispice.C('out', 0, Cl)
    return cct


def cfg3_oscillator(L1=1e-6, C1=1e-9):
    cct = _new("cfg3 nonlinear oscillator")
    cct.L1 = eispice.L('a', 0, L1)
    cct.C1 = eispice.C('a', 0, C1)
    cct.B1 = eispice.B('a', 0, eispice.Current, '-1m*v(a) + 1m*v(a)^3')
    cct.I1 = eispice.I('a', 0, 0, eispice.Pulse(0, '1m', 0, '1n', '1n', '10n', 1))
    return cct


# (builder, analysis, known-answer checks)
#   analysis: ('op',) or ('tran', tstep, tstop[, tmax])
CASES = {
    "inductor": (inductor, ('tran', '0.5n', '100n'),
                 [('v', 1, 1.333333e+01, '1.85e-8'), ('v', 1, -2.000000e+01, '1.05e-8')]),
    "capacitor": (capacitor, ('tran', '0.5n', '100n'),
                  [('i', 'Vx', 1.333333e+01, '1.85e-8'), ('i', 'Vx', -2.000000e+01, '1.05e-8')]),
    "resistor": (resistor, ('op',), [('v', 1, 1.26e+01, 0), ('i', 'Vx', -3.88889e-02, 0)]),
    "realc": (realc, ('tran', '0.5n', '100n'),
              [('i', 'Vx', -4.832985, '1.85e-8'), ('i', 'Vx', -0.2574886, '1.05e-8')]),
    "b_current": (b_current, ('op',),
                  [('v', 1, 7, 0), ('v', 2, 10.44121563, 0), ('i', 'Vx', -0.7, 0)]),
    "b_voltage_time": (b_voltage_time, ('tran', '0.5n', '15n'),
                       [('v', 2, 9.486671194e-01, '3n'), ('v', 2, 3.088685702e-01, '10.5n')]),
    "b_capacitor": (b_capacitor, ('tran', '0.5n', '100n'),
                    [('i', 'Vx', -1.810333469e+02, '12n'), ('i', 'Vx', -3.770187057e2, '71n')]),
    "i_pulse": (i_pulse, ('tran', '1n', '20n'),
                [('i', 'Vx', 4, '5n'), ('i', 'Vx', 8, '15n')]),
    "v_pulse": (v_pulse, ('tran', '1n', '20n'),
                [('v', 1, 4, '5n'), ('v', 1, 8, '15n')]),
    "vi_table": (vi_table, ('op',), [('i', 'VIx', 10, 0)]),
    "vi_table_neg": (lambda: vi_table(-5), ('op',), [('i', 'VIx', -2, 0)]),
    "vccs": (vccs, ('op',), [('i', 'Vy', -6.4, 0)]),
    "vcvs": (vcvs, ('op',), [('v', 2, 8.64, 0)]),
    "cccs": (cccs, ('op',), [('i', 'Vy', 5.6, 0)]),
    "ccvs": (ccvs, ('op',), [('v', 2, -8.64, 0)]),
    "tline": (tline, ('tran', '0.01n', '10n'),
              [('v', 'vo', 0.3726765, '2.6n'), ('i', 'Vx', -0.008381298823, '4.62n')]),
    "tline_lossy": (tline_lossy, ('tran', '0.01n', '10n'), []),
    "diode": (diode, ('tran', '0.5u', '10u'),
              [('i', 'Vx', -7.326775239e-02, '4.6u'), ('i', 'Vx', -1.781444540e-01, '6.4u')]),
    "w_pwl": (w_pwl, ('tran', '0.5n', '100n'),
              [('v', 1, 8.815789474e+00, '25n'), ('v', 1, -2.000000000e+01, '75n')]),
    "w_pwc": (w_pwc, ('tran', '0.5n', '100n'),
              [('v', 1, 1.148836888e+01, '25n'), ('v', 1, -2.000000000e+01, '75n')]),
    "w_sffm": (w_sffm, ('tran', '0.5n', '100n'),
               [('v', 1, -2.630296891e+00, '25n'), ('v', 1, 4.631842886e+00, '75n')]),
    "w_exp": (w_exp, ('tran', '0.5n', '100n'),
              [('v', 1, 3.999818400e+00, '25n'), ('v', 1, 1.818267660e-04, '75n')]),
    "w_pulse": (w_pulse, ('tran', '0.5n', '100n'),
                [('v', 1, 4, '25n'), ('v', 1, 8, '75n')]),
    "w_gauss": (w_gauss, ('tran', '0.5n', '100n'),
                [('v', 1, 1.517639357e-01, '25n'), ('v', 1, 3.148220583e+00, '65n')]),
    "w_sin": (w_sin, ('tran', '0.5n', '100n'),
              [('v', 1, -6.561833244e-04, '25n'), ('v', 1, 1.910792387e-04, '75n')]),
    "subckt_nested": (subckt_nested, ('op',), [('i', 'Vx', -10.0 / 300.0, 0)]),
    "circuit_basic": (circuit_basic, ('tran', '1n', '2n'), [('v', 1, 10, '1n')]),
    "circuit_tran": (circuit_tran, ('tran', '0.1n', '1n', '0.5n'), [('v', 1, 1, '0.5n')]),
    "cfg1_rc": (cfg1_rc, ('tran', '0.01n', '10n'), []),
    "cfg2_rectifier": (cfg2_rectifier, ('tran', '10n', '3u'), []),
    "cfg3_oscillator": (cfg3_oscillator, ('tran', '1n', '1u'), []),
}


# ---- IBIS circuits (tables: tests/golden/ibis_tables.npz, generated from the
# reference's own fixture module/ibis_test.py by tests/golden/make_golden.py) --
import json as _json
import os as _os

from eispice_b200 import ibis_device as _ibis

_IBIS = None


def ibis_fixture():
    """(models, pins): plain dictionaries in the layout ibis_device expects."""
    global _IBIS
    if _IBIS is None:
        path = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                             "golden", "ibis_tables.npz")
        z = np.load(path)
        meta = _json.loads(str(z["meta"]))
        models = {}
        for name, mm in meta["models"].items():
            m = {"model_type": mm["model_type"],
                 "voltage_range": mm["voltage_range"], "c_comp": mm["c_comp"]}
            for tab in mm["tables"]:
                m[tab] = {sp: z["%s/%s/%s" % (name, tab, sp)]
                          for sp in ("typ", "min", "max")}
            for tab in mm["k_tables"]:
                m[tab] = {dr: {sp: z["%s/%s/%s/%s" % (name, tab, dr, sp)]
                               for sp in ("typ", "min", "max")}
                          for dr in ("rising", "falling")}
            models[name] = m
        # V-t waveforms, fixtures and ramp data (tests/golden/make_ibis_vt.py)
        vt = np.load(_os.path.join(_os.path.dirname(path), "ibis_vt.npz"))
        vmeta = _json.loads(str(vt["meta"]))
        for name, vm in vmeta.items():
            for direction, waves in vm["waveforms"].items():
                models[name][direction + "_waveform"] = [
                    dict(w, data={sp: vt["%s/%s/%d/%s" % (name, direction, k, sp)]
                                  for sp in ("typ", "min", "max")})
                    for k, w in enumerate(waves)]
            if vm["ramp"]:
                models[name]["ramp"] = vm["ramp"]
        _IBIS = (models, meta["pins"])
    return _IBIS


def ibis_pin(pin_name, node, speed="typ", direction="rising"):
    models, pins = ibis_fixture()
    p = pins[str(pin_name)]
    return _ibis.Pin(models[p["model"]], p, node, speed, direction)


def ibis_link(driver_pin='2', sections=None, Rt='33.2', speed="typ",
              direction="rising"):
    """Driver pin -> Vmeas -> Rt -> line -> receiver pin '1'
    (reference module/ibis.py:61-70).  ``sections``: None = one lossless
    T(50, 2n); otherwise [(Z0, Td, loss), ...] cascaded (cfg5)."""
    cct = _new('IBIS link')
    cct.Driver = ibis_pin(driver_pin, 'vsx', speed, direction)
    cct.Vmeas = eispice.V('vsx', 'vs', 0)
    cct.Rt = eispice.R('vs', 'vi', Rt)
    if sections is None:
        cct.Tg = eispice.T('vi', 0, 'vo', 0, 50, '2n')
    else:
        prev = 'vi'
        for k, (z0, td, loss) in enumerate(sections):
            nxt = 'vo' if k == len(sections) - 1 else 'n%d' % (k + 1)
            setattr(cct, 'T%d' % (k + 1), eispice.T(prev, 0, nxt, 0, z0, td, loss))
            prev = nxt
    cct.Receiver = ibis_pin('1', 'vo', speed, direction)
    return cct


def cfg4_ibis():
    return ibis_link('2')


def ibis_ramp():                     # reference module/ibis.py:73-84
    return ibis_link('4')


def cfg5_ibis_lossy():
    return ibis_link('2', sections=[(50, 0.25e-9, 0.05)] * 8)


IBIS_CASES = {
    "cfg4_ibis": (cfg4_ibis, ('tran', '0.01n', '15n'),
                  [('v', 'vo', 2.942970279e-01, '4.7n'),
                   ('i', 'Vmeas', 2.200965422e-02, '7.1n')]),
    "ibis_ramp": (ibis_ramp, ('tran', '0.01n', '15n'),
                  [('v', 'vo', -2.416134420e-01, '1.7n'),
                   ('i', 'Vmeas', 1.521229286e-02, '4.9n')]),
    "cfg5_ibis_lossy": (cfg5_ibis_lossy, ('tran', '0.01n', '15n'), []),
}

# result columns kept in tests/golden/ref_waveforms.npz for the big cases
GOLDEN_COLUMNS = {
    "cfg4_ibis": ['v(vsx)', 'i(Vmeas)', 'v(vi)', 'v(vo)', 'v(0@die)', 'v(3@die)'],
    "ibis_ramp": ['v(vsx)', 'i(Vmeas)', 'v(vi)', 'v(vo)', 'v(0@die)', 'v(3@die)'],
    "cfg5_ibis_lossy": ['v(vsx)', 'i(Vmeas)', 'v(vi)', 'v(vo)', 'v(n4)', 'v(3@die)'],
}


# ---- circuits without committed reference waveforms: pinned through the live
# reference in the build container (tests/test_oracle_golden.py) and compared
# oracle <-> GPU on the GPU box ------------------------------------------------
def vi_spline_ta():
    """VI device with a cubic-spline V-I table and a time-dependent multiplier
    (math/piecewise.c:108-181, devices/vicurve.c:85-106), behind an RC."""
    cct = _new("VI spline + TA")
    vi = np.array([[-2, -0.08], [-1, -0.02], [0, 0.0], [0.5, 0.004], [1, 0.012],
                   [2, 0.05], [3, 0.06]], dtype=float)
    ta = np.array([[0, 0.0], [2e-9, 0.0], [3e-9, 1.0], [7e-9, 1.0], [8e-9, 0.3], [20e-9, 0.3]])
    cct.Vx = eispice.V(1, 0, 0, eispice.Pulse(0, 2.5, '1n', '1n', '1n', '6n', '20n'))
    cct.Rs = eispice.R(1, 2, 25)
    cct.Cl = eispice.C(2, 0, '2p')
    cct.VIx = eispice.VI(2, 0, eispice.PWC(vi), eispice.PWL(ta))
    return cct


def pwl_current_into_rlc():
    """PWL current source into a parallel RLC (source_i.c, inductor.c, capacitor.c)."""
    cct = _new("PWL I into RLC")
    cct.Ix = eispice.I(1, 0, 0, eispice.PWL([[0, 0], ['1n', '1m'], ['3n', '-1m'], ['6n', 0]]))
    cct.Rp = eispice.R(1, 0, 200)
    cct.Lp = eispice.L(1, 0, '20n')
    cct.Cp = eispice.C(1, 0, '5p')
    return cct


EXTRA_CASES = {
    "vi_spline_ta": (vi_spline_ta, ('tran', '0.05n', '12n'), []),
    "pwl_current_into_rlc": (pwl_current_into_rlc, ('tran', '0.02n', '10n'), []),
}


def all_cases():
    out = dict(CASES)
    out.update(IBIS_CASES)
    return out
