"""The Python front end mirrors the reference's module/*.py call shapes: unit
strings (module/units.py doctests :123-130,160-163,187-188), sub-circuit name
mangling (module/subckt.py:60-66), the diode built from R + two B sources with
the reference's equation strings (module/device.py:805-829), waveform argument
slots (simulator.h:75).  No GPU needed."""
import math

import numpy as np
import pytest

import eispice_b200 as eispice
from eispice_b200 import subckt, units


def test_units_doctests_of_the_reference():
    assert units.float(10.0) == 10.0 and units.float(10) == 10.0
    assert units.float('10nF') == 1e-08
    assert units.float('12nF/2nH') == 6.0
    # value * multiplier, like the reference (so '0.3m' is 0.3 * 1e-3, one ulp off 3e-4)
    assert np.array_equal(units.floatList2D([['0.3m', 1.2], ['3n', '5u']]),
                          np.array([[0.3 * 1e-3, 1.2], [3 * 1e-9, 5 * 1e-6]]))
    assert np.array_equal(units.floatList1D(['0.3m', 1.2, '3n', '5u']),
                          np.array([0.3 * 1e-3, 1.2, 3 * 1e-9, 5 * 1e-6]))
    # 'M' is mega here (and milli inside B equations, libs/calculon/tokenizer.re:154-164)
    assert units.float('100M') == 1e8 and units.float('1Meg') == 1e6 and units.float('2k') == 2e3
    assert units.float('.5u') == 5e-7 and units.float('1e-3') == 1e-3


def test_subcircuit_names_are_mangled_like_the_reference():
    subckt.reset_counter()

    class Div(eispice.Subckt):
        def __init__(self, a, b):
            self.R1 = eispice.R(a, self.node('mid'), 100)
            self.R2 = eispice.R(self.node('mid'), b, 100)

    cct = eispice.Circuit("x")
    cct.X1 = Div('in', 0)
    cct.X2 = Div('in', 0)
    refs = [d[1] for d in cct.netlist()]
    assert refs == ['0#R1', '0#R2', '1#R1', '1#R2']
    assert cct.netlist()[0][3] == '0@mid' and cct.netlist()[2][3] == '1@mid'
    with pytest.raises(ValueError):
        cct._add('0#R1', eispice.R(1, 0, 1))            # refdes listed twice


def test_diode_is_two_b_sources_and_a_resistor():
    subckt.reset_counter()
    cct = eispice.Circuit("d")
    cct.D1 = eispice.D('a', 'k', IS='2.52n', RS=0.568, N=1.752, CJO='4p', M=0.4, TT='20n')
    nl = cct.netlist()
    assert [d[0] for d in nl] == ['R', 'B', 'B']
    assert nl[1][4] == 'i' and nl[2][4] == 'c'
    assert nl[1][5].startswith("(if(v(0@r,k)>=") and "exp(v(0@r,k)/" in nl[1][5]
    assert "^4.000000e-01" in nl[2][5]                  # parameters printed with %e
    assert abs(nl[0][4] - 0.568) < 1e-15


def test_waveform_argument_slots_and_defaults():
    p = eispice.Pulse(0, 1, '1n', '0.1n')
    assert p.kind == 'p' and p.args7()[:4] == [0.0, 1.0, 1 * 1e-9, 0.1 * 1e-9]
    assert all(math.isinf(x) for x in p.args7()[4:])    # unset -> HUGE_VAL, replaced per run
    s = eispice.Sin(0, 5, '1Meg')
    assert s.kind == 's' and s.args7()[:3] == [0.0, 5.0, 1e6] and len(s.args7()) == 7
    t = eispice.PWL([['12n', 3], ['2n', 4]])
    assert t.kind == 'l' and t.pw[0, 0] == 2e-9        # sorted by x like the reference
    with pytest.raises(TypeError):
        eispice.Pulse(1)


def test_controlled_sources_are_b_equations():
    g = eispice.G(2, 0, 1, 0, 2)
    assert g.kind == 'B' and g.type == 'i' and g.equation == "v(1,0)*2.000000e+00"
    h = eispice.H(2, 0, 'Vx', 3.2)
    assert h.type == 'v' and h.equation == "i(Vx)*3.200000e+00"
