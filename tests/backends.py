"""Test-side runners: replay one netlist on the CPU oracle (oracle/eis_oracle.c)
and, when it was built, on the UNMODIFIED reference engine (oracle/_ref).

Test infrastructure only.  A netlist is the list of tuples produced by
``eispice_b200.Circuit.netlist()``:

  ('R'|'C'|'L', refdes, p, n, value)
  ('V'|'I', refdes, p, n, 'v'|'i', dc, wave)   wave: None | (kind, [7 args]) | ('l'|'c', table)
  ('B', refdes, p, n, type, equation)
  ('T', refdes, pl, nl, pr, nr, Z0, Td, loss)
  ('VI', refdes, p, n, [(type, table), ...], [(type, table), ...] | None)
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libeis_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libeispice_ref.so")

_dp = C.POINTER(C.c_double)


def build_oracle():
    """(Re)build the oracle; also builds oracle/_ref when /root/reference exists."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True,
                   stdout=subprocess.DEVNULL)


def _b(s):
    return str(s).encode()


class EoStats(C.Structure):
    _fields_ = [("accepted", C.c_long), ("rejected_lte", C.c_long),
                ("rejected_nc", C.c_long), ("factorizations", C.c_long),
                ("op_factorizations", C.c_long)]


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        lib = C.CDLL(ORACLE_SO)
        lib.eoNew.restype = C.c_void_p
        lib.eoDestroy.argtypes = [C.c_void_p]
        s3 = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]
        for f in ("eoAddResistor", "eoAddCapacitor", "eoAddInductor"):
            getattr(lib, f).argtypes = s3 + [C.c_double]
        lib.eoAddSource.argtypes = s3 + [C.c_char, C.c_double, C.c_char,
                                         C.c_void_p, C.c_void_p, C.c_int]
        lib.eoAddNonlinearSource.argtypes = s3 + [C.c_char, C.c_char_p]
        lib.eoAddTLine.argtypes = [C.c_void_p] + [C.c_char_p] * 5 + [C.c_double] * 3
        lib.eoAddVICurve.argtypes = s3 + [C.c_void_p, C.c_int, C.c_char,
                                          C.c_void_p, C.c_int, C.c_char]
        lib.eoSetParam.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_double]
        lib.eoSetPivots.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        lib.eoRunTransient.argtypes = [C.c_void_p, C.c_double, C.c_double,
                                       C.c_double, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        lib.eoRunOperatingPoint.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        lib.eoFree.argtypes = [C.c_void_p]
        lib.eoNumUnknowns.argtypes = [C.c_void_p]
        lib.eoNumNonzeros.argtypes = [C.c_void_p]
        lib.eoGetPattern.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.eoGetStats.argtypes = [C.c_void_p, C.c_void_p]
        lib.eoCalcEval.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_double, C.c_void_p, C.c_void_p]
        lib.eoSetAttemptLog.argtypes = [C.c_void_p, C.c_int]
        lib.eoGetAttemptLog.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.eoGetLinearizeLog.argtypes = [C.c_void_p, C.c_void_p]
        lib.eoGetPivots.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.eoSetStampWatch.argtypes = [C.c_void_p, C.c_int]
        lib.eoGetStampWatch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _oracle = lib
    return _oracle


_ref = None


def ref_available():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.simulatorNew.restype = C.c_void_p
        lib.simulatorNew.argtypes = [C.c_void_p]
        lib.simulatorDestroy.argtypes = [C.c_void_p]
        s3 = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]
        for f in ("simulatorAddResistor", "simulatorAddCapacitor",
                  "simulatorAddInductor"):
            getattr(lib, f).argtypes = s3 + [C.c_void_p]
        lib.simulatorAddSource.argtypes = s3 + [C.c_char, C.c_void_p, C.c_char,
                                                C.c_void_p]
        lib.simulatorAddNonlinearSource.argtypes = s3 + [C.c_char, C.c_char_p]
        lib.simulatorAddTLine.argtypes = [C.c_void_p] + [C.c_char_p] * 5 + [C.c_void_p] * 3
        lib.simulatorAddVICurve.argtypes = s3 + [C.c_void_p, C.c_void_p, C.c_char,
                                                 C.c_void_p, C.c_void_p, C.c_char]
        lib.simulatorRunTransient.argtypes = [C.c_void_p, C.c_double, C.c_double,
                                              C.c_double, C.c_int, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p]
        lib.simulatorRunOperatingPoint.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        lib.ref_shim_quiet.argtypes = [C.c_int]
        lib.ref_calc_eval.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_double, C.c_void_p, C.c_void_p]
        lib.ref_probe_capture.argtypes = [C.c_int]
        lib.ref_probe_count.argtypes = [C.c_void_p]
        lib.ref_probe_sequence.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        _ref = lib
    return _ref


def _table(t):
    return np.ascontiguousarray(np.asarray(t, dtype=np.float64))


class OracleSim:
    """One instance of a netlist on the oracle."""

    def __init__(self, netlist, vi_set=0):
        self.lib = oracle_lib()
        self.h = C.c_void_p(self.lib.eoNew())
        lib, h = self.lib, self.h
        for d in netlist:
            k = d[0]
            if k in ("R", "C", "L"):
                fn = {"R": lib.eoAddResistor, "C": lib.eoAddCapacitor,
                      "L": lib.eoAddInductor}[k]
                rc = fn(h, _b(d[1]), _b(d[2]), _b(d[3]), float(d[4]))
            elif k in ("V", "I"):
                wave = d[6]
                if wave is None:
                    rc = lib.eoAddSource(h, _b(d[1]), _b(d[2]), _b(d[3]), _b(k.lower()),
                                         float(d[5]), b"\0", None, None, 0)
                elif wave[0] in ("l", "c"):
                    t = _table(wave[1])
                    rc = lib.eoAddSource(h, _b(d[1]), _b(d[2]), _b(d[3]), _b(k.lower()),
                                         float(d[5]), _b(wave[0]), None,
                                         t.ctypes.data, len(t))
                else:
                    a = np.array(wave[1], dtype=np.float64)
                    rc = lib.eoAddSource(h, _b(d[1]), _b(d[2]), _b(d[3]), _b(k.lower()),
                                         float(d[5]), _b(wave[0]), a.ctypes.data,
                                         None, 0)
            elif k == "B":
                rc = lib.eoAddNonlinearSource(h, _b(d[1]), _b(d[2]), _b(d[3]),
                                              _b(d[4]), _b(d[5]))
            elif k == "T":
                rc = lib.eoAddTLine(h, _b(d[1]), _b(d[2]), _b(d[3]), _b(d[4]),
                                    _b(d[5]), float(d[6]), float(d[7]), float(d[8]))
            elif k == "VI":
                vis, tas = d[4], d[5]
                vt, vtab = vis[vi_set if len(vis) > 1 else 0]
                vtab = _table(vtab)
                if tas:
                    tt, ttab = tas[vi_set if len(tas) > 1 else 0]
                    ttab = _table(ttab)
                    rc = lib.eoAddVICurve(h, _b(d[1]), _b(d[2]), _b(d[3]),
                                          vtab.ctypes.data, len(vtab), _b(vt),
                                          ttab.ctypes.data, len(ttab), _b(tt))
                else:
                    rc = lib.eoAddVICurve(h, _b(d[1]), _b(d[2]), _b(d[3]),
                                          vtab.ctypes.data, len(vtab), _b(vt),
                                          None, 0, b"l")
            else:
                raise ValueError("unknown device kind %r" % (k,))
            if rc != 0:
                raise RuntimeError("oracle rejected device %r" % (d[1],))

    def set_param(self, refdes, name, value):
        if self.lib.eoSetParam(self.h, _b(refdes), _b(name), float(value)):
            raise KeyError("%s.%s" % (refdes, name))

    def set_pivots(self, phase, pc, pr):
        pc = np.ascontiguousarray(pc, dtype=np.int32)
        pr = np.ascontiguousarray(pr, dtype=np.int32)
        self.lib.eoSetPivots(self.h, phase, pc.ctypes.data, pr.ctypes.data, len(pc))

    def _collect(self, rc, data, names, npts, nvar):
        if rc != 0:
            raise RuntimeError("oracle run failed")
        arr = np.ctypeslib.as_array(data, shape=(npts.value, nvar.value)).copy()
        var = [names[i].decode() for i in range(nvar.value)]
        self.lib.eoFree(data)
        self.lib.eoFree(names)
        return arr, var

    def tran(self, tstep, tstop, tmax=0.0, restart=True):
        data, names = _dp(), C.POINTER(C.c_char_p)()
        npts, nvar = C.c_int(), C.c_int()
        rc = self.lib.eoRunTransient(self.h, tstep, tstop, tmax, int(restart),
                                     C.byref(data), C.byref(names),
                                     C.byref(npts), C.byref(nvar))
        return self._collect(rc, data, names, npts, nvar)

    def op(self):
        data, names = _dp(), C.POINTER(C.c_char_p)()
        npts, nvar = C.c_int(), C.c_int()
        rc = self.lib.eoRunOperatingPoint(self.h, C.byref(data), C.byref(names),
                                          C.byref(npts), C.byref(nvar))
        return self._collect(rc, data, names, npts, nvar)

    def log_attempts(self, on=True):
        """Record (time, order, outcome, solves) of every transient attempt."""
        self.lib.eoSetAttemptLog(self.h, 1 if on else 0)

    def attempt_log(self):
        """(times[n], flags[n]) of the last transient; flags = order |
        outcome << 4 | solves << 8, outcome 0 accepted / 1 LTE / 2 non-convergence."""
        t, f = _dp(), C.POINTER(C.c_int)()
        n = self.lib.eoGetAttemptLog(self.h, C.byref(t), C.byref(f))
        if n == 0:
            return np.zeros(0), np.zeros(0, dtype=np.int32)
        return (np.ctypeslib.as_array(t, shape=(n,)).copy(),
                np.ctypeslib.as_array(f, shape=(n,)).copy())

    def linearize_log(self):
        """One word per linearize pass (operating point first): bit k = the
        convergence tests of the k-th nonlinear device passed."""
        m = C.POINTER(C.c_int)()
        n = self.lib.eoGetLinearizeLog(self.h, C.byref(m))
        if n == 0:
            return np.zeros(0, dtype=np.int32)
        return np.ctypeslib.as_array(m, shape=(n,)).copy()

    def first_pivots(self, phase):
        """(pc, pr) of the first factorisation of a phase in the runs so far."""
        n = self.lib.eoNumUnknowns(self.h)
        pc = np.zeros(n, dtype=np.int32)
        pr = np.zeros(n, dtype=np.int32)
        if self.lib.eoGetPivots(self.h, phase, pc.ctypes.data, pr.ctypes.data) < 0:
            return None
        return pc, pr

    def watch_stamps(self, on=True):
        self.lib.eoSetStampWatch(self.h, 1 if on else 0)

    def stamp_changes(self, phase):
        """(A[N, N] bool, B[N] bool): what differed between two consecutive solves
        of one Newton loop in the runs since watch_stamps() (None: nothing ran)."""
        a, b = C.POINTER(C.c_char)(), C.POINTER(C.c_char)()
        n = self.lib.eoGetStampWatch(self.h, phase, C.byref(a), C.byref(b))
        if n == 0:
            return None
        A = np.frombuffer(C.string_at(a, n * n), dtype=np.uint8).reshape(n, n).astype(bool)
        B = np.frombuffer(C.string_at(b, n), dtype=np.uint8).astype(bool)
        return A, B

    def stats(self):
        st = EoStats()
        self.lib.eoGetStats(self.h, C.byref(st))
        return {f: getattr(st, f) for f, _ in EoStats._fields_}

    def pattern(self):
        n = self.lib.eoNumUnknowns(self.h)
        nnz = self.lib.eoNumNonzeros(self.h)
        ri = np.zeros(nnz, dtype=np.int32)
        cs = np.zeros(n + 1, dtype=np.int32)
        self.lib.eoGetPattern(self.h, ri.ctypes.data, cs.ctypes.data)
        return n, ri, cs

    def close(self):
        if self.h:
            self.lib.eoDestroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


class RefSim:
    """One instance of a netlist on the unmodified reference engine.  The
    reference borrows every parameter pointer, so the ctypes
    objects are kept alive on ``self``."""

    def __init__(self, netlist, vi_set=0, quiet=True):
        self.lib = ref_lib()
        self.lib.ref_shim_quiet(1 if quiet else 0)
        self.keep = []
        self.params = {}
        self.h = C.c_void_p(self.lib.simulatorNew(None))
        lib, h = self.lib, self.h
        for d in netlist:
            k = d[0]
            if k in ("R", "C", "L"):
                fn = {"R": lib.simulatorAddResistor, "C": lib.simulatorAddCapacitor,
                      "L": lib.simulatorAddInductor}[k]
                rc = fn(h, _b(d[1]), _b(d[2]), _b(d[3]), self._dbl(d[1], k, d[4]))
            elif k in ("V", "I"):
                wave = d[6]
                dc = self._dbl(d[1], "DC", d[5])
                if wave is None:
                    rc = lib.simulatorAddSource(h, _b(d[1]), _b(d[2]), _b(d[3]),
                                                _b(k.lower()), dc, b"\0", None)
                else:
                    args = (C.c_void_p * 7)()
                    if wave[0] in ("l", "c"):
                        t = _table(wave[1])
                        tp = C.c_void_p(t.ctypes.data)
                        ln = C.c_int(len(t))
                        self.keep += [t, tp, ln]
                        args[0] = C.cast(C.pointer(tp), C.c_void_p)
                        args[1] = C.cast(C.pointer(ln), C.c_void_p)
                    else:
                        for i, v in enumerate(wave[1]):
                            args[i] = C.cast(self._dbl(d[1], "A%d" % i, v), C.c_void_p)
                    self.keep.append(args)
                    rc = lib.simulatorAddSource(h, _b(d[1]), _b(d[2]), _b(d[3]),
                                                _b(k.lower()), dc, _b(wave[0]), args)
            elif k == "B":
                eq = C.create_string_buffer(_b(d[5]))
                self.keep.append(eq)
                rc = lib.simulatorAddNonlinearSource(h, _b(d[1]), _b(d[2]), _b(d[3]),
                                                     _b(d[4]), eq)
            elif k == "T":
                rc = lib.simulatorAddTLine(h, _b(d[1]), _b(d[2]), _b(d[3]), _b(d[4]),
                                           _b(d[5]), self._dbl(d[1], "Z0", d[6]),
                                           self._dbl(d[1], "Td", d[7]),
                                           self._dbl(d[1], "loss", d[8]))
            elif k == "VI":
                vis, tas = d[4], d[5]
                vt, vtab = vis[vi_set if len(vis) > 1 else 0]
                vp, vl = self._tab(vtab)
                if tas:
                    tt, ttab = tas[vi_set if len(tas) > 1 else 0]
                    tp, tl = self._tab(ttab)
                    rc = lib.simulatorAddVICurve(h, _b(d[1]), _b(d[2]), _b(d[3]),
                                                 vp, vl, _b(vt), tp, tl, _b(tt))
                else:
                    rc = lib.simulatorAddVICurve(h, _b(d[1]), _b(d[2]), _b(d[3]),
                                                 vp, vl, _b(vt), None, None, b"l")
            else:
                raise ValueError("unknown device kind %r" % (k,))
            if rc != 0:
                raise RuntimeError("reference rejected device %r" % (d[1],))

    def _dbl(self, refdes, name, v):
        p = C.c_double(float(v))
        self.keep.append(p)
        self.params[(refdes, name)] = p
        return C.cast(C.pointer(p), C.c_void_p)

    def _tab(self, t):
        t = _table(t)
        tp = C.c_void_p(t.ctypes.data)
        ln = C.c_int(len(t))
        self.keep += [t, tp, ln]
        return C.cast(C.pointer(tp), C.c_void_p), C.cast(C.pointer(ln), C.c_void_p)

    def set_param(self, refdes, name, value):
        self.params[(refdes, name)].value = float(value)

    def _collect(self, rc, data, names, npts, nvar):
        if rc != 0:
            raise RuntimeError("reference run failed")
        arr = np.ctypeslib.as_array(data, shape=(npts.value, nvar.value)).copy()
        var = [names[i].decode() for i in range(nvar.value)]
        return arr, var  # (leaks the two mallocs; libc free is not worth binding)

    def tran(self, tstep, tstop, tmax=0.0, restart=True):
        data, names = _dp(), C.POINTER(C.c_char_p)()
        npts, nvar = C.c_int(), C.c_int()
        rc = self.lib.simulatorRunTransient(self.h, tstep, tstop, tmax, int(restart),
                                            C.byref(data), C.byref(names),
                                            C.byref(npts), C.byref(nvar))
        return self._collect(rc, data, names, npts, nvar)

    def op(self):
        data, names = _dp(), C.POINTER(C.c_char_p)()
        npts, nvar = C.c_int(), C.c_int()
        rc = self.lib.simulatorRunOperatingPoint(self.h, C.byref(data), C.byref(names),
                                                 C.byref(npts), C.byref(nvar))
        return self._collect(rc, data, names, npts, nvar)


def superlu_sequences(netlist, vi_set, tstep, tstop, max_records=100000):
    """Run one instance on the unmodified reference and return the pivot
    sequences of every full factorisation SuperLU did (oracle/ref_probe.c):
    (pc[k][N], pr[k][N]); k = 0 is the operating point, k = 1 the first
    transient attempt."""
    lib = ref_lib()
    sim = RefSim(netlist, vi_set)
    lib.ref_probe_capture(max_records)
    try:
        sim.tran(tstep, tstop)
    except RuntimeError:
        pass          # the reference aborted; what it factorised until then was captured
    finally:
        n = C.c_int(0)
        cnt = lib.ref_probe_count(C.byref(n))
        lib.ref_probe_capture(0)
    pc = np.zeros((cnt, n.value), dtype=np.int32)
    pr = np.zeros((cnt, n.value), dtype=np.int32)
    for k in range(cnt):
        lib.ref_probe_sequence(k, pc[k].ctypes.data, pr[k].ctypes.data)
    return pc, pr


def eval_expr(lib_kind, expr, names, values, min_div=1e-15):
    """(rc, value, derivs) of a calculon expression on 'oracle' or 'ref'."""
    n = len(names)
    arr = (C.c_char_p * max(n, 1))(*[_b(x) for x in names])
    vals = (C.c_double * max(n, 1))(*values)
    val = C.c_double(0.0)
    der = (C.c_double * max(n, 1))()
    if lib_kind == "oracle":
        rc = oracle_lib().eoCalcEval(_b(expr), n, arr, vals, min_div, C.byref(val), der)
    else:
        rc = ref_lib().ref_calc_eval(_b(expr), n, arr, vals, min_div, C.byref(val), der)
    return rc, val.value, list(der)[:n]
