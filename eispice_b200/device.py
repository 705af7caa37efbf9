"""Device classes, mirroring the reference's ``module/device.py`` (R C L V I B
VI G E F H T D RealC) on top of the batched engine.

A device is a plain record (kind, nodes, parameters).  It is attached to a
:class:`eispice_b200.Circuit` by attribute assignment exactly like in eispice.
Scalar parameters keep the reference's member names (``R1.R``, ``T1.Z0``, ...:
reference module/simulatormodule.c:660-662,870,1273,1330,1387) and are the
names used for per-instance sweeps in ``Circuit.tran_batch``.
"""
import math

from . import subckt, units

Current = "i"
Voltage = "v"
Capacitor = "c"
Time = "time"
GND = "0"


class Device:
    """Base record: ``kind`` selects the simulator_cuda ``Add*`` call."""
    kind = None
    params = ()

    def _nodes(self, *nodes):
        self.node = [str(n) for n in nodes]


class R(Device):
    """Resistor(pNode, nNode, R)"""
    kind = "R"
    params = ("R",)

    def __init__(self, pNode, nNode, R):
        self._nodes(pNode, nNode)
        self.R = units.float(R)


class C(Device):
    """Capacitor(pNode, nNode, C)"""
    kind = "C"
    params = ("C",)

    def __init__(self, pNode, nNode, C):
        self._nodes(pNode, nNode)
        self.C = units.float(C)


class L(Device):
    """Inductor(pNode, nNode, L)"""
    kind = "L"
    params = ("L",)

    def __init__(self, pNode, nNode, L):
        self._nodes(pNode, nNode)
        self.L = units.float(L)


class _Source(Device):
    params = ("DC",)

    def __init__(self, pNode, nNode, dcValue=0.0, wave=None):
        self._nodes(pNode, nNode)
        self.DC = units.float(dcValue)
        self.wave = wave


class V(_Source):
    """VoltageSource(pNode, nNode, dcValue=0, wave=None)"""
    kind = "V"
    type = "v"


class I(_Source):  # noqa: E742 - eispice name
    """CurrentSource(pNode, nNode, dcValue=0, wave=None)"""
    kind = "I"
    type = "i"


class B(Device):
    """Behavioural source B(pNode, nNode, type, equation); type is Current,
    Voltage or Capacitor; the equation is a calculon expression
    (reference libs/calculon)."""
    kind = "B"

    def __init__(self, pNode, nNode, type, equation):
        self._nodes(pNode, nNode)
        self.type = str(type)
        self.equation = str(equation)


class G(B):
    """Voltage-controlled current source (reference module/device.py:489-513)"""

    def __init__(self, pNode, nNode, pControlNode, nControlNode, value):
        B.__init__(self, pNode, nNode, Current, "v(%s,%s)*%e" % (
            str(pControlNode), str(nControlNode), units.float(value)))


class E(B):
    """Voltage-controlled voltage source (reference module/device.py:515-539)"""

    def __init__(self, pNode, nNode, pControlNode, nControlNode, value):
        B.__init__(self, pNode, nNode, Voltage, "v(%s,%s)*%e" % (
            str(pControlNode), str(nControlNode), units.float(value)))


class F(B):
    """Current-controlled current source (reference module/device.py:541-569)"""

    def __init__(self, pNode, nNode, controlDevice, value):
        B.__init__(self, pNode, nNode, Current, "i(%s)*%e" % (
            str(controlDevice), units.float(value)))


class H(B):
    """Current-controlled voltage source (reference module/device.py:571-599)"""

    def __init__(self, pNode, nNode, controlDevice, value):
        B.__init__(self, pNode, nNode, Voltage, "i(%s)*%e" % (
            str(controlDevice), units.float(value)))


class VI(Device):
    """V-I table device VI(pNode, nNode, vi, ta=None): current = ta(t)*vi(v).

    ``vi`` / ``ta`` are PWL/PWC tables, or lists of tables (one per table set,
    e.g. IBIS corner) selected per instance with the ``set`` parameter."""
    kind = "VI"

    def __init__(self, pNode, nNode, vi, ta=None):
        self._nodes(pNode, nNode)
        self.VI = vi
        self.TA = ta


class T(Device):
    """Lossless/lossy delay line T(pL, nL, pR, nR, Z0, Td, loss=None)"""
    kind = "T"
    params = ("Z0", "Td", "loss")

    def __init__(self, pNodeLeft, nNodeLeft, pNodeRight, nNodeRight, Z0, Td,
                 loss=None):
        self._nodes(pNodeLeft, nNodeLeft, pNodeRight, nNodeRight)
        self.Z0 = units.float(Z0)
        self.Td = units.float(Td)
        self.loss = math.inf if loss is None else units.float(loss)


class RealC(subckt.Subckt):
    """Capacitor with ESL and ESR (reference module/device.py:140-167)"""

    def __init__(self, pNode, nNode, cap, ESL, ESR):
        self.C = C(pNode, self.node("n0"), cap)
        self.R = R(self.node("n0"), self.node("n1"), ESR)
        self.L = L(self.node("n1"), nNode, ESL)


class D(subckt.Subckt):
    """spice3f5-style junction diode built from R + two B sources, exactly as
    the reference composes it (module/device.py:746-829): the parameters are
    printed into the equations with ``%e``."""

    def __init__(self, pNode, nNode, area=1.0, IS=1.0e-14, RS=0, N=1, TT=0,
                 CJO=0, VJ=1, M=0.5, EG=1.11, XTI=3.0, KF=0, AF=1, FC=0.5,
                 BV=1e100, IBV=1e-3, TNOM=27):
        pNode, nNode = str(pNode), str(nNode)
        area, IS, RS, N = (units.float(x) for x in (area, IS, RS, N))
        TT, CJO, VJ, M = (units.float(x) for x in (TT, CJO, VJ, M))
        BV, IBV, TNOM = (units.float(x) for x in (BV, IBV, TNOM))

        if RS != 0:
            self.Rs = R(pNode, self.node("r"), area * RS)
            pNode = self.node("r")

        k = 1.3806503e-23
        q = 1.60217646e-19
        Vt = ((k * (TNOM + 273.15)) / q)

        Is = ("(if(v(%s,%s)>=%e)*(%e*(exp(v(%s,%s)/%e)-1)))" %
              (pNode, nNode, -BV, area * IS, pNode, nNode, N * Vt))
        Ib = ("(if(v(%s,%s)<%e)*(%e*(exp((-%e-v(%s,%s))/%e)-1)))" %
              (pNode, nNode, -BV, area * IS, BV, pNode, nNode, Vt))
        self.Isb = B(pNode, nNode, Current, Is + "+" + Ib)

        Cd = "(%e*(exp(v(%s,%s)/%e)))" % (area * TT * IS / N * Vt, pNode,
                                           nNode, N * Vt)
        Cj1 = ("(if(v(%s,%s)<%e)*(%e/((1-v(%s,%s)/%e)^%e)))" %
               (pNode, nNode, FC * VJ, area * CJO, pNode, nNode, VJ, M))
        Cj2 = ("(if(v(%s,%s)>=%e)*%e*(1-%e+%e*v(%s,%s)))" %
               (pNode, nNode, FC * VJ, (area * CJO / ((1 - FC) ** (M + 1))),
                FC * (1 + M), (M / VJ), pNode, nNode))
        self.Cjd = B(pNode, nNode, Capacitor, Cd + "+" + Cj1 + "+" + Cj2)
