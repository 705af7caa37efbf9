"""Source waveforms, mirroring the reference's ``module/waveform.py`` and the
waveform types of ``module/simulatormodule.c`` (members V1/V2/Td/... at
simulatormodule.c:147-154,209-217,274-283,341-350,406-413).

Unset optional arguments are ``inf`` (HUGE_VAL) and are replaced per run by the
engine exactly as the reference does (libs/simulator/math/waveform.c:342-413).
"""
import math

import numpy as np

from . import units

_UNSET = math.inf


class _Wave:
    kind = None
    _names = ()
    _required = 2

    def __init__(self, *args):
        vals = list(units.floatList1D(args))
        if len(vals) < self._required or len(vals) > len(self._names):
            raise TypeError("%s takes %d to %d arguments" % (
                type(self).__name__, self._required, len(self._names)))
        for i, name in enumerate(self._names):
            setattr(self, name, float(vals[i]) if i < len(vals) else _UNSET)

    def args7(self):
        """The 7-slot argument vector of simulatorAddSource (simulator.h:75)."""
        vals = [getattr(self, n) for n in self._names]
        return vals + [_UNSET] * (7 - len(vals))

    def param_index(self, name):
        return self._names.index(name)


class Pulse(_Wave):
    """Pulse(V1, V2, [Td, Tr, Tf, PW, Per])"""
    kind = "p"
    _names = ("V1", "V2", "Td", "Tr", "Tf", "PW", "Per")


class Gauss(_Wave):
    """Gauss(V1, V2, [Td, Tr, Tf, PW, Per]) -- pulse with Gaussian edges"""
    kind = "g"
    _names = ("V1", "V2", "Td", "Tr", "Tf", "PW", "Per")


class Sin(_Wave):
    """Sin(Vo, Va, [Fc, Td, DF])"""
    kind = "s"
    _names = ("Vo", "Va", "Fc", "Td", "DF")


class Exp(_Wave):
    """Exp(V1, V2, [Td1, Tau1, Td2, Tau2])"""
    kind = "e"
    _names = ("V1", "V2", "Td1", "Tau1", "Td2", "Tau2")


class SFFM(_Wave):
    """SFFM(Vo, Va, [Fc, MDI, Fs])"""
    kind = "f"
    _names = ("Vo", "Va", "Fc", "MDI", "Fs")


class _Table:
    kind = None

    def __init__(self, data):
        data = np.asarray(units.floatList2D(data), dtype=np.float64)
        if data.ndim != 2 or data.shape[1] != 2:
            raise ValueError("table must be [n][2]")
        # sorted by x for the simulator (reference module/waveform.py:69)
        self.pw = np.ascontiguousarray(data[data[:, 0].argsort(), ])
        self.type = self.kind


class PWL(_Table):
    """Piece-wise linear table [[x, y], ...]"""
    kind = "l"


class PWC(_Table):
    """Piece-wise natural-cubic-spline table [[x, y], ...]"""
    kind = "c"
