"""``Circuit``: host-side mirror of the reference's ``module/circuit.py`` /
``Circuit_`` (module/simulatormodule.c:1502-1632) on the batched CUDA engine.

``tran()`` / ``op()`` keep the reference's call shape and result attributes
(``results``, ``variables``, ``t``, ``v[...]``, ``i[...]``, ``check_v``,
``check_i``); they run ONE instance through the same CUDA path the batch uses.
``tran_batch()`` is the additive batch entry of SURVEY.md 8b: per-instance
parameter arrays in, probe waveforms on the ``tstep`` grid out.

There is no CPU path: without the CUDA extension / a GPU these calls raise.
"""
import numpy as np

from . import subckt, units, waveform
from .device import Device

Current = "i"
Voltage = "v"
Time = "time"
GND = "0"


class _Interp:
    """Linear interpolation of one result column (reference circuit.py:48-67)."""

    def __init__(self, t, y):
        self.time = t if len(t) > 1 else None
        self.all = y

    def __call__(self, time):
        if self.time is None:
            return float(self.all[0]) if np.ndim(self.all) else float(self.all)
        return float(np.interp(units.float(time), self.time, self.all))


class _StrKeyDict(dict):
    def __getitem__(self, key):
        return dict.__getitem__(self, str(key))


def _wave_entry(wave):
    if wave is None:
        return None
    if isinstance(wave, (waveform.PWL, waveform.PWC)):
        return (wave.kind, wave.pw)
    return (wave.kind, [float(x) for x in wave.args7()])


def _tables(obj):
    if obj is None:
        return None
    items = obj if isinstance(obj, (list, tuple)) else [obj]
    return [(t.kind, t.pw) for t in items]


class Circuit:
    def __init__(self, title="eispice"):
        d = self.__dict__
        d["title"] = title
        d["_devices"] = []          # (refdes, Device) in insertion order
        d["_engine"] = None
        d["results"] = None
        d["variables"] = None

    # -- netlist construction (reference circuit.py:129-138, simulatormodule.c:1590-1632)
    def __setattr__(self, name, value):
        if isinstance(value, subckt.Subckt):
            for sub_name, sub_value in value:
                self._add(sub_name, sub_value)
            self.__dict__[name] = value
        elif isinstance(value, Device):
            self._add(name, value)
        else:
            self.__dict__[name] = value

    def _add(self, refdes, dev):
        if self._engine is not None:
            raise RuntimeError("Can't add a device to a simulator that has run.")
        if any(r == refdes for r, _ in self._devices):
            raise ValueError("%s is listed twice" % refdes)
        self._devices.append((refdes, dev))
        self.__dict__[refdes] = dev

    def netlist(self):
        """The flat device list as tuples (see tests/backends.py for the layout)."""
        out = []
        for refdes, d in self._devices:
            n = d.node
            if d.kind in ("R", "C", "L"):
                out.append((d.kind, refdes, n[0], n[1], getattr(d, d.kind)))
            elif d.kind in ("V", "I"):
                out.append((d.kind, refdes, n[0], n[1], d.type, d.DC, _wave_entry(d.wave)))
            elif d.kind == "B":
                out.append(("B", refdes, n[0], n[1], d.type, d.equation))
            elif d.kind == "T":
                out.append(("T", refdes, n[0], n[1], n[2], n[3], d.Z0, d.Td, d.loss))
            elif d.kind == "VI":
                out.append(("VI", refdes, n[0], n[1], _tables(d.VI), _tables(d.TA)))
            else:
                raise TypeError("unsupported device %r" % (d,))
        return out

    # -- analyses ---------------------------------------------------------
    def _get_engine(self, num_instances=1):
        from . import engine
        eng = self._engine
        if eng is None or eng.num_instances != num_instances:
            if eng is not None:
                eng.close()
            eng = engine.Engine(self.netlist(), num_instances)
            self.__dict__["_engine"] = eng
        else:
            eng.refresh_scalars(self.netlist())
        return eng

    def tran(self, tstep, tstop, tmax=0.0, restart=False):
        """Transient analysis of this one circuit (reference circuit.py:192-217)."""
        eng = self._get_engine(1)
        data, names = eng.tran_single(units.float(tstep), units.float(tstop),
                                      units.float(tmax), bool(restart))
        self.__dict__["results"] = data
        self.__dict__["variables"] = names
        self._results()

    def op(self):
        """Operating point (reference circuit.py:220-236)."""
        eng = self._get_engine(1)
        data, names = eng.op_single()
        self.__dict__["results"] = data
        self.__dict__["variables"] = names
        self._results()

    def tran_batch(self, tstep, tstop, tmax=0.0, params=None, probes=None,
                   num_instances=None, **options):
        """Batched transient over independent instances of this topology.

        params -- {"Refdes.attr": array[M]} per-instance scalars
                  (``R C L DC Z0 Td loss``, waveform members such as ``V2``,
                  and ``set`` = table-set index of a VI device)
        probes -- result names such as ``"v(out)"`` / ``"i(Vx)"``
        Returns an ``engine.BatchResult`` (``.data[M, ngrid, nprobes]``,
        ``.grid``, ``.status[M]``, ``.stats``).
        """
        params = params or {}
        if num_instances is None:
            sizes = {len(np.atleast_1d(v)) for v in params.values()}
            num_instances = max(sizes) if sizes else 1
        eng = self._get_engine(int(num_instances))
        return eng.tran_batch(units.float(tstep), units.float(tstop),
                              units.float(tmax), params, probes or [], **options)

    def op_batch(self, params=None, probes=None, num_instances=None):
        """Operating point of every instance of a sweep: ([M, nprobes], status[M])."""
        params = params or {}
        if num_instances is None:
            sizes = {len(np.atleast_1d(v)) for v in params.values()}
            num_instances = max(sizes) if sizes else 1
        eng = self._get_engine(int(num_instances))
        return eng.op_batch(list(probes or []), params)

    # -- result access (reference circuit.py:140-190,241-298) ---------------
    def _results(self):
        res = self.results
        d = self.__dict__
        d["t"] = res[:, 0]
        d["i"] = {}
        d["v"] = _StrKeyDict()
        for n, name in enumerate(self.variables):
            if name[0] == "i":
                self.i[name[2:-1]] = _Interp(self.t, res[:, n])
            if name[0] == "v":
                self.v[name[2:-1]] = _Interp(self.t, res[:, n])

    def current_array(self, *variables):
        idx = [self.variables.index("time")]
        idx += [self.variables.index("i(%s)" % str(v)) for v in variables]
        return self.results.take(idx, 1)

    def voltage_array(self, *variables):
        idx = [self.variables.index("time")]
        idx += [self.variables.index("v(%s)" % str(v)) for v in variables]
        return self.results.take(idx, 1)

    @staticmethod
    def _check(name, value, sim_value):
        sign = 1 if value >= 0 else -1
        lo = value * (1.0 - sign * 1e-4)
        hi = value * (1.0 + sign * 1e-4)
        if sim_value < lo or sim_value > hi:
            print("%s FAIL: %.9e != %.9e" % (name, value, sim_value))
            return False
        return True

    def check_v(self, name, value, time=0.0):
        return self._check(str(name), units.float(value),
                           self.v[str(name)](units.float(time)))

    def check_i(self, name, value, time=0.0):
        return self._check(str(name), units.float(value),
                           self.i[str(name)](units.float(time)))
