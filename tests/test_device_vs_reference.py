"""Device code against the UNMODIFIED reference, function by function, on the CPU.

The numerical-integration and break-point modules of the kernel are
re-organised relative to libs/simulator/math (rings stored by age, one shared
time ring, the negative-ring-index quirk spelled out, the angle test decided
from its arguments).  These tests compile exactly those device functions of
eispice_b200/csrc/simcu_device.cuh for the HOST (g++, no contraction) and drive
them side by side with the reference's own integrator.c / checkbreak.c from
oracle/_ref/libeispice_ref.so (built from the reference's sources by
oracle/build_ref.sh) through the call pattern of core/simulator.c:142-233 -
attempts that advance, attempts that step back after a reject, order changes.
Every output must be bit-identical.

Test infrastructure only: a host build of device functions is not a CPU path
of the product (the engine has none).  Skipped where oracle/_ref is absent.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import backends

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEVICE = os.path.join(ROOT, "eispice_b200", "csrc", "simcu_device.cuh")
PROGRAM = os.path.join(ROOT, "eispice_b200", "csrc", "simcu_program.h")

pytestmark = pytest.mark.skipif(not backends.ref_available(), reason="needs oracle/_ref (the reference build)")


class Control(C.Structure):
    """control_ of libs/simulator/include/control.h:35-56"""
    _fields_ = [("itl1", C.c_int), ("itl4", C.c_int), ("reltol", C.c_double), ("vntol", C.c_double),
                ("captol", C.c_double), ("abstol", C.c_double), ("chgtol", C.c_double),
                ("trtol", C.c_double), ("minstep", C.c_double), ("gmin", C.c_double),
                ("maxorder", C.c_int), ("luLibrary", C.c_int), ("maxAngleA", C.c_double),
                ("maxAngleV", C.c_double), ("tstop", C.c_double), ("tstep", C.c_double),
                ("integratorOrder", C.c_int), ("time", C.c_double)]


def _control(tstop=15e-9, tstep=1e-11):
    c = Control()
    c.itl1, c.itl4, c.maxorder = 100, 10, 2
    c.reltol, c.vntol, c.abstol, c.captol = 1e-3, 1e-6, 1e-12, 1e-18
    c.chgtol, c.trtol, c.gmin = 1e-14, 1.0, 1e-15
    c.minstep = tstep * 5e-5
    c.maxAngleA = c.maxAngleV = np.pi / 3
    c.tstop, c.tstep = tstop, tstep
    c.integratorOrder, c.time = 1, 0.0
    return c


def _host_device_lib(tmp_path):
    dev = open(DEVICE).read()
    prog = open(PROGRAM).read()
    integ = dev[dev.index("DEV double DD1("):dev.index("#else\n#define RING(c, sb, k, j)")] + "#endif\n"
    wave = dev[dev.index("DEV double wfArg("):dev.index("/* ------------------------------------------------------------------------ *\n * Behavioural sources:")]
    conv = dev[dev.index("DEV int clIsLinear("):dev.index("/* ------------------------------------------------------------------------ *\n * Integrator:")]
    brk = dev[dev.index("DEVFN int cbAngleBreak("):
              dev.index("#else\n#define cbIsBreakClocked cbIsBreak\nDEV void cbInitialize")]
    code = """
#include <cmath>
#include <cstring>
using std::isnan;
%s
#define DEV static inline
#define DEVFN static
#define SIMCU_JIT 1
#define SIMCU_SHAREDCB 1
#define SIMCU_UNROLL
#define HUGE_VAL (__builtin_huge_val())
#define MaxAbs(x, y) ((fabs(x) > fabs(y)) ? fabs(x) : fabs(y))
#define MUL_RN(a, b) ((a) * (b))
#define ADD_RN(a, b) ((a) + (b))
#define W_TSTOP(c) ((c).a.ctl.tstop)
enum { SIMCU_ERR_NAN = 3 };
static double simcuAtan2(double y, double x) { return (y == 0.0 && x > 0.0) ? y : atan2(y, x); }
struct Inst {
	SimcuArgs a;
	double time; int order; int err;
	mutable double sv[64]; mutable int iv[8];
	mutable double itT[4], itH[3]; mutable int itN; mutable bool itAdv;
	mutable double cbT0, cbT1, cbDt1; mutable int cbN; mutable bool cbAdv;
	double par[8];
	double P(int i) const { return par[i]; }
	int tableSet(int) const { return 0; }
	double &S(int i) const { return sv[i]; }
	int &IS(int i) const { return iv[i]; }
};
%s
%s
%s
/* table waveforms ('l' / 'c') are not exercised here: stubs for the code to compile */
#define W_TSTEP(c) ((c).a.ctl.tstep)
struct Table { const double *p; int len; int type; };
static Table tableOf(const SimcuArgs &, int) { Table t = {0, 0, 0}; return t; }
static void pwCalcValue(const Table &, int &, double, double &y, double &d) { y = 0; d = 0; }
static double pwGetNextX(const Table &, int &, double) { return HUGE_VAL; }
%s
extern "C" {
void *devNew(double tstop, double reltol, double chgtol, double trtol, double gmin)
{
	Inst *c = new Inst();
	std::memset(c, 0, sizeof *c);
	c->a.ctl.tstop = tstop; c->a.ctl.reltol = reltol; c->a.ctl.chgtol = chgtol;
	c->a.ctl.trtol = trtol; c->a.ctl.gmin = gmin;
	return c;
}
void devFree(void *h) { delete (Inst *)h; }
void devItInit(void *h, double ic, double ydtdx) { itInitialize(*(Inst *)h, 0, 0, ic, ydtdx); }
int devItIntegrate(void *h, double time, int order, double x0, double ydtdx, double *dydx0, double *y0)
{
	Inst &c = *(Inst *)h;
	c.time = time; c.order = order; c.err = 0;
	itBegin(c);
	itIntegrate(c, 0, 0, x0, ydtdx, *dydx0, *y0);
	return c.err;
}
int devItNextStep(void *h, double x0, double ydtdx, double abstol, double *step)
{
	Inst &c = *(Inst *)h;
	c.err = 0;
	*step = itNextStep(c, 0, 0, x0, ydtdx, abstol);
	return c.err;
}
void devCbInit(void *h, double ic)
{
	Inst &c = *(Inst *)h;
	cbInitialize(c, 30, 0, ic); cbInitialize(c, 40, 1, ic);
	c.cbT0 = 0.0; c.cbT1 = 0.0; c.cbDt1 = 0.0; c.cbN = 0; c.cbAdv = false;      /* as at the claim */
}
/* the per-device clock (a T-line's second test) and the instance's shared clock */
int devCbIsBreak(void *h, double time, double x, double maxAngle, int *clocked)
{
	Inst &c = *(Inst *)h;
	c.time = time;
	cbBegin(c);
	*clocked = cbIsBreakClocked(c, 40, 1, x, maxAngle);
	return cbIsBreak(c, 30, 0, x, maxAngle);
}
void devWave(void *h, int type, const double *args, double time, double tstep, double tstop,
		double *value, double *next)
{
	Inst &c = *(Inst *)h;
	int d[SIMCU_DEV_STRIDE] = {0};
	d[DF_WTYPE] = type; d[DF_PARGS] = 0;
	for (int k = 0; k < 7; k++) c.par[k] = args[k];
	c.a.ctl.tstep = tstep; c.a.ctl.tstop = tstop;
	c.time = time;
	*value = wfValue(c, d, 0, 0);
	wfNextStep(c, d, 0, *next);
}
void devClInit(void *h, double ic) { Inst &c = *(Inst *)h; c.S(50) = ic; c.IS(4) = 0; }
int devClIsLinear(void *h, double tol, double V, double calcedV) { return clIsLinear(*(Inst *)h, 50, 4, tol, V, calcedV); }
}
""" % (prog, integ, brk, conv, wave)
    cpp, so = tmp_path / "dev.cpp", tmp_path / "dev.so"
    cpp.write_text(code)
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-w", "-shared", "-fPIC", "-o", str(so), str(cpp)],
                   check=True)
    lib = C.CDLL(str(so))
    lib.devNew.restype = C.c_void_p
    lib.devNew.argtypes = [C.c_double] * 5
    lib.devFree.argtypes = [C.c_void_p]
    lib.devItInit.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lib.devItIntegrate.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    lib.devItNextStep.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]
    lib.devCbInit.argtypes = [C.c_void_p, C.c_double]
    lib.devCbIsBreak.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]
    lib.devWave.argtypes = [C.c_void_p, C.c_int, C.c_void_p] + [C.c_double] * 3 + [C.c_void_p, C.c_void_p]
    lib.devClInit.argtypes = [C.c_void_p, C.c_double]
    lib.devClIsLinear.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    return lib


def _reference():
    ref = backends.ref_lib()
    ref.integratorNew.restype = C.c_void_p
    ref.integratorNew.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_char]
    ref.integratorInitialize.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
    ref.integratorIntegrate.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    ref.integratorNextStep.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
    ref.checkbreakNew.restype = C.c_void_p
    ref.checkbreakNew.argtypes = [C.c_void_p, C.c_void_p, C.c_char]
    ref.checkbreakInitialize.argtypes = [C.c_void_p, C.c_double]
    ref.checkbreakIsBreak.argtypes = [C.c_void_p, C.c_double]
    ref.waveformNew.restype = C.c_void_p
    ref.waveformNew.argtypes = [C.c_void_p, C.c_void_p, C.c_char, C.c_void_p, C.c_void_p]
    ref.waveformInitialize.argtypes = [C.c_void_p]
    ref.waveformCalcValue.argtypes = [C.c_void_p, C.c_void_p]
    ref.waveformNextStep.argtypes = [C.c_void_p, C.c_void_p]
    ref.checklinearNew.restype = C.c_void_p
    ref.checklinearNew.argtypes = [C.c_void_p, C.c_void_p, C.c_char]
    ref.checklinearInitialize.argtypes = [C.c_void_p, C.c_double]
    ref.checklinearIsLinear.argtypes = [C.c_void_p, C.c_double, C.c_double]
    return ref


def _bits(x):
    return np.float64(x).tobytes()


def _attempt_times(rng, n, tstep):
    """Attempt times as core/simulator.c:156-218 produces them: forward after an
    accept, a shorter step from the last accepted time after a reject."""
    t, accepted = [], []
    last_ok, now = 0.0, 0.0
    step = tstep / 100
    for _ in range(n):
        now = last_ok + step
        ok = rng.random() > 0.25
        t.append(now)
        accepted.append(ok)
        if ok:
            last_ok = now
            step = step * rng.choice([0.1, 1.0, 2.0])
        else:
            step = step * rng.choice([0.125, 0.5, 0.9])
    return t, accepted


@pytest.mark.parametrize("units,ydtdx_kind", [("V", "const"), ("A", "const"), ("F", "varying")])
def test_integrator_equals_the_references_integrator(tmp_path, units, ydtdx_kind):
    """math/integrator.c:61-163 incl. the negative ring index while the counter is
    below the look-back (:90-96) and the float sqrtf (:99): companion model (G,
    Ieq) and LTE step of the device code, bit for bit, on random runs."""
    dev, ref = _host_device_lib(tmp_path), _reference()
    rng = np.random.default_rng({"V": 1, "A": 2, "F": 3}[units])
    compared = 0
    for trial in range(60):
        ctl = _control()
        ydtdx = C.c_double(10.0 ** rng.uniform(-13, -8))
        abstol = {"V": ctl.vntol, "A": ctl.abstol, "F": ctl.captol}[units]
        h = dev.devNew(ctl.tstop, ctl.reltol, ctl.chgtol, ctl.trtol, ctl.gmin)
        r = ref.integratorNew(None, C.byref(ctl), C.byref(ydtdx), units.encode())
        assert r
        ic = float(rng.normal())
        # devLoad initialises with 0, devInitStep with the operating point
        dev.devItInit(h, 0.0, ydtdx.value)
        assert ref.integratorInitialize(r, ic, C.byref(ydtdx)) == 0
        dev.devItInit(h, ic, ydtdx.value)
        times, accepted = _attempt_times(rng, 120, ctl.tstep)
        x_ok = ic
        order = 2                                        # core/simulator.c:133
        for tm, ok in zip(times, accepted):
            if ydtdx_kind == "varying":
                ydtdx.value = ydtdx.value * float(rng.uniform(0.9, 1.1))
            ctl.time, ctl.integratorOrder = tm, order
            d0, y0 = C.c_double(), C.c_double()
            assert ref.integratorIntegrate(r, x_ok, C.byref(d0), C.byref(y0)) == 0
            d1, y1 = C.c_double(), C.c_double()
            assert dev.devItIntegrate(h, tm, order, x_ok, ydtdx.value, C.byref(d1), C.byref(y1)) == 0
            assert _bits(d0.value) == _bits(d1.value) and _bits(y0.value) == _bits(y1.value), (trial, tm, order)
            x_new = x_ok + float(rng.normal()) * 10.0 ** rng.uniform(-6, -1) * (rng.random() > 0.2)
            s0, s1 = C.c_double(), C.c_double()
            rc0 = ref.integratorNextStep(r, x_new, C.byref(s0))
            rc1 = dev.devItNextStep(h, x_new, ydtdx.value, abstol, C.byref(s1))
            assert (rc0 != 0) == (rc1 != 0)
            if rc0 == 0:
                assert _bits(s0.value) == _bits(s1.value), (trial, tm, order, s0.value, s1.value)
                compared += 1
            if ok:
                x_ok = x_new
                order = min(order + 1, 2)
            else:
                order = 1 if rng.random() < 0.5 else max(order - 1, 1)
        dev.devFree(h)
    assert compared > 5000


def test_break_point_test_equals_the_references_checkbreak(tmp_path):
    """math/checkbreak.c:43-67 against the device code - the per-device version (state
    by age, arguments instead of angles) and the one with the instance's shared
    clock (cbBegin / cbIsBreakClocked): same decision at every call."""
    dev, ref = _host_device_lib(tmp_path), _reference()
    rng = np.random.default_rng(17)
    calls = breaks = 0
    for trial in range(80):
        ctl = _control()
        h = dev.devNew(ctl.tstop, ctl.reltol, ctl.chgtol, ctl.trtol, ctl.gmin)
        r = ref.checkbreakNew(None, C.byref(ctl), b"V")
        assert r
        assert ref.checkbreakInitialize(r, 0.0) == 0
        dev.devCbInit(h, 0.0)
        times, _ = _attempt_times(rng, 200, ctl.tstep)
        kind = trial % 4
        x = 0.0
        for k, tm in enumerate(times):
            if kind == 0:
                x += float(rng.normal()) * 1e-2
            elif kind == 1:
                x = 1.8 if (k // 25) % 2 else 0.0                      # pulse plateaus and edges
            elif kind == 2:
                x += float(rng.normal()) * 1e-12 * (rng.random() > 0.5)   # angles near the threshold
            else:
                x = 3.3 * np.sin(2e9 * tm)
            ctl.time = tm
            want = ref.checkbreakIsBreak(r, x)
            clocked = C.c_int()
            got = dev.devCbIsBreak(h, tm, x, ctl.maxAngleV, C.byref(clocked))
            assert want == got == clocked.value, (trial, k, tm, x)
            calls += 1
            breaks += got
        dev.devFree(h)
    assert calls == 16000 and breaks > 200


@pytest.mark.parametrize("units", ["V", "A", "F"])
def test_convergence_test_equals_the_references_checklinear(tmp_path, units):
    """math/checklinear.c:39-91: two tolerance tests at reltol and the a -> b -> c
    state that survives across timesteps - the decision that ends a Newton loop."""
    dev, ref = _host_device_lib(tmp_path), _reference()
    rng = np.random.default_rng(ord(units))
    ctl = _control()
    tol = {"V": ctl.vntol, "A": ctl.abstol, "F": ctl.captol}[units]
    scale = {"V": 1.0, "A": 1e-2, "F": 1e-12}[units]
    passes = calls = 0
    for trial in range(200):
        h = dev.devNew(ctl.tstop, ctl.reltol, ctl.chgtol, ctl.trtol, ctl.gmin)
        r = ref.checklinearNew(None, C.byref(ctl), units.encode())
        ic = float(rng.normal()) * scale
        assert ref.checklinearInitialize(r, ic) == 0
        dev.devClInit(h, ic)
        v = ic
        for k in range(60):
            u = rng.random()
            if u < 0.5:                      # converging: changes around the tolerance
                v = v * (1 + float(rng.normal()) * 10.0 ** rng.uniform(-6, -2))
            elif u < 0.6:
                v = float(rng.normal()) * scale
            calced = v * (1 + float(rng.normal()) * 10.0 ** rng.uniform(-6, -2)) + float(rng.normal()) * tol
            want = ref.checklinearIsLinear(r, v, calced)
            got = dev.devClIsLinear(h, tol, v, calced)
            assert want == got, (trial, k, v, calced)
            passes += got
            calls += 1
        dev.devFree(h)
    assert calls == 12000 and 500 < passes < calls - 500


@pytest.mark.parametrize("kind", ["p", "g", "s", "e", "f"])
def test_source_waveforms_equal_the_references_waveform(tmp_path, kind):
    """math/waveform.c:169-413: Pulse, Gauss, Sin, Exp and SFFM - value and the
    look-ahead to the next corner - with every combination of omitted arguments
    (HUGE_VAL), whose defaults depend on tstep / tstop of the run (:342-413)."""
    dev, ref = _host_device_lib(tmp_path), _reference()
    rng = np.random.default_rng(ord(kind))
    nargs = {"p": 7, "g": 7, "s": 5, "e": 6, "f": 5}[kind]
    optional = {"p": (2, 4, 5, 6), "g": (2, 4, 5, 6), "s": (2, 3, 4), "e": (2, 3, 4, 5), "f": (2, 3, 4)}[kind]
    compared = 0
    for trial in range(150):
        tstep = 10.0 ** rng.uniform(-12, -9)
        tstop = tstep * float(rng.integers(50, 2000))
        ctl = _control(tstop, tstep)
        a = np.full(7, np.inf)
        a[0], a[1] = rng.normal(), rng.normal() * 3
        span = tstop
        if kind in "pg":
            a[2:7] = [rng.uniform(0, 0.2) * span, rng.uniform(0.01, 0.1) * span, rng.uniform(0.01, 0.1) * span,
                      rng.uniform(0.05, 0.3) * span, rng.uniform(0.5, 1.2) * span]
        elif kind == "s":
            a[2:5] = [rng.uniform(1, 10) / span, rng.uniform(0, 0.3) * span, rng.uniform(0, 3) / span]
        elif kind == "e":
            td1 = rng.uniform(0, 0.3) * span
            a[2:6] = [td1, rng.uniform(0.02, 0.2) * span, td1 + rng.uniform(0.1, 0.5) * span,
                      rng.uniform(0.02, 0.2) * span]
        else:
            a[2:5] = [rng.uniform(1, 10) / span, rng.uniform(0, 5), rng.uniform(0.5, 5) / span]
        for k in optional:
            if rng.random() < 0.35:
                a[k] = np.inf                              # omitted: the reference's default applies
        cells = [C.c_double(float(v)) for v in a]
        argv = (C.c_void_p * 7)(*[C.cast(C.pointer(c), C.c_void_p) for c in cells])
        dc = C.c_void_p()
        r = ref.waveformNew(None, C.byref(ctl), kind.encode(), argv, C.byref(dc))
        assert r
        assert ref.waveformInitialize(r) == 0
        h = dev.devNew(tstop, ctl.reltol, ctl.chgtol, ctl.trtol, ctl.gmin)
        args = np.ascontiguousarray(a)
        for tm in np.concatenate(([0.0], np.sort(rng.uniform(0, tstop, 60)), [tstop])):
            ctl.time = float(tm)
            v0, n0 = C.c_double(), C.c_double(-7.0)
            assert ref.waveformCalcValue(r, C.byref(v0)) == 0
            assert ref.waveformNextStep(r, C.byref(n0)) == 0
            v1, n1 = C.c_double(), C.c_double(-7.0)
            dev.devWave(h, ord(kind), args.ctypes.data, float(tm), tstep, tstop, C.byref(v1), C.byref(n1))
            assert _bits(v0.value) == _bits(v1.value), (kind, trial, tm, a.tolist(), v0.value, v1.value)
            assert _bits(n0.value) == _bits(n1.value), (kind, trial, tm, a.tolist(), n0.value, n1.value)
            compared += 1
        dev.devFree(h)
        _ = nargs
    assert compared == 150 * 62


def _host_table_lib(tmp_path):
    """pwSetIndex / pwCalcValueFn / pwGetNextX of the device library on packed tables."""
    dev = open(DEVICE).read()
    lookups = dev[dev.index("/* piecewise.c:44-67.  The cached index only shortens the walk;"):
                  dev.index("/* ------------------------------------------------------------------------ *\n * Break-point test:")]
    code = """
#include <cmath>
#define DEV static inline
#define DEVFN static
#define HUGE_VAL (__builtin_huge_val())
struct Table { const double *p; int len; int type; };
struct TabEntry { double x, y, m, xn; };
static TabEntry tabLoad(const Table &t, int i)
{
	TabEntry e; const double *q = t.p + 4 * (size_t)i;
	e.x = q[0]; e.y = q[1]; e.m = q[2]; e.xn = q[3];
	return e;
}
%s
extern "C" {
void devTable(const double *packed, int len, int type, int *index, double x0, double *y, double *dydx, double *nextx)
{
	Table t; t.p = packed; t.len = len; t.type = type;
	int i = *index;
	pwCalcValue(t, i, x0, *y, *dydx);
	int j = *index;
	*nextx = pwGetNextX(t, j, x0);
	*index = i;
}
}
""" % lookups
    cpp, so = tmp_path / "tab.cpp", tmp_path / "tab.so"
    cpp.write_text(code)
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-w", "-shared", "-fPIC", "-o", str(so), str(cpp)],
                   check=True)
    lib = C.CDLL(str(so))
    lib.devTable.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


@pytest.mark.parametrize("kind", ["l", "c"])
def test_table_lookups_equal_the_references_piecewise(tmp_path, kind):
    """math/piecewise.c:44-67,96-263: linear and natural-cubic-spline tables as the
    host packs them (simcuPackTable: slopes / second derivatives) and the device
    looks them up (index walk from a cached index, value, slope, next corner) -
    against piecewiseNew / Initialize / CalcValue / GetNextX of the reference."""
    from eispice_b200 import engine
    L = engine.lib()
    L.simcuPackTable.argtypes = [C.c_void_p, C.c_int, C.c_char, C.c_void_p]
    dev, ref = _host_table_lib(tmp_path), _reference()
    ref.piecewiseNew.restype = C.c_void_p
    ref.piecewiseNew.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_char]
    ref.piecewiseInitialize.argtypes = [C.c_void_p]
    ref.piecewiseCalcValue.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    ref.piecewiseGetNextX.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    rng = np.random.default_rng(ord(kind))
    compared = 0
    for trial in range(120):
        n = int(rng.integers(2 if kind == "l" else 3, 40))
        x = np.cumsum(rng.uniform(0.01, 1.0, n)) * 10.0 ** rng.uniform(-10, 0) - rng.uniform(0, 2)
        y = np.cumsum(rng.normal(size=n)) * 10.0 ** rng.uniform(-3, 1)
        xy = np.ascontiguousarray(np.stack([x, y], axis=1))
        packed = np.zeros((n, 4))
        assert L.simcuPackTable(xy.ctypes.data, n, kind.encode(), packed.ctypes.data) == 0
        data = C.c_void_p(xy.ctypes.data)
        length = C.c_int(n)
        r = ref.piecewiseNew(None, C.byref(data), C.byref(length), kind.encode())
        assert r
        assert ref.piecewiseInitialize(r) == 0
        span = x[-1] - x[0]
        i_ref, i_dev = C.c_int(0), C.c_int(0)
        # a walk like a Newton loop's: small moves, occasional jumps, beyond both ends
        q = x[0] + rng.uniform(-0.2, 1.2) * span
        for k in range(80):
            u = rng.random()
            if u < 0.7:
                q += rng.normal() * 0.05 * span
            elif u < 0.85:
                q = x[0] + rng.uniform(-0.3, 1.3) * span
            else:
                q = float(x[int(rng.integers(0, n))])              # exactly on a table point
            y0, d0, nx0 = C.c_double(), C.c_double(), C.c_double()
            j = C.c_int(i_ref.value)
            assert ref.piecewiseCalcValue(r, C.byref(i_ref), q, C.byref(y0), C.byref(d0)) == 0
            assert ref.piecewiseGetNextX(r, C.byref(j), q, C.byref(nx0)) == 0
            y1, d1, nx1 = C.c_double(), C.c_double(), C.c_double()
            dev.devTable(packed.ctypes.data, n, ord(kind), C.byref(i_dev), q, C.byref(y1), C.byref(d1), C.byref(nx1))
            assert i_ref.value == i_dev.value, (kind, trial, k, q)
            assert _bits(y0.value) == _bits(y1.value) and _bits(d0.value) == _bits(d1.value), (kind, trial, k, q, y0.value, y1.value)
            assert _bits(nx0.value) == _bits(nx1.value), (kind, trial, k, q, nx0.value, nx1.value)
            compared += 1
    assert compared == 120 * 80


def test_tline_history_lookup_equals_the_references_history_interp(tmp_path):
    """core/history.c:36-59 + math/history_interp.c:42-92 on the reference's own
    linked list of accepted records against the kernel's ring: galloping + binary
    search from a hint (tlineLocateFn) and the interpolation / extrapolation of a
    record column (tlineInterp).  Same value, bit for bit; the same lookups fail.
    (At or past the newest record the reference itself crashes, see below.)"""
    dev_src = open(DEVICE).read()
    body = dev_src[dev_src.index("DEV int tlineWrap(int p, int R)"):dev_src.index("#ifdef SIMCU_JIT\n/* First pass over the T-lines")]
    code = """
#include <cstddef>
#define DEV static inline
#define DEVFN static
struct int4 { int x, y, z, w; };
%s
extern "C" int lookup(const double *times, const double *vals, int nh, int R, int count, int *hint, double tp,
		int col, double *out)
{
	const int4 r = tlineLocateFn(times, R, count, *hint, tp);
	if (!r.w) return -1;
	*hint = r.x;
	*out = tlineInterp(vals + (size_t)r.y * nh, vals + (size_t)r.z * nh, col, times[r.y], times[r.z], tp);
	return 0;
}
""" % body
    cpp, so = tmp_path / "hist.cpp", tmp_path / "hist.so"
    cpp.write_text(code)
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-w", "-shared", "-fPIC", "-o", str(so), str(cpp)],
                   check=True)
    dev = C.CDLL(str(so))
    dev.lookup.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_int,
                           C.c_void_p]
    ref = _reference()
    ref.listNew.restype = C.c_void_p
    ref.listNew.argtypes = [C.c_void_p]
    ref.listAdd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    ref.historyNew.restype = C.c_void_p
    ref.historyNew.argtypes = [C.c_double, C.c_void_p, C.c_int, C.c_uint]
    ref.historyInterpNew.restype = C.c_void_p
    ref.historyInterpNew.argtypes = [C.c_void_p, C.c_void_p]
    ref.historyInterpInitialize.argtypes = [C.c_void_p]
    ref.historyInterpSetTime.argtypes = [C.c_void_p, C.c_double]
    ref.historyInterpGetData.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    rng = np.random.default_rng(33)
    same = failed = past = 0
    nh = 5
    for trial in range(60):
        count = int(rng.integers(1, 120))
        R = count + int(rng.integers(0, 8))                 # nothing has left the ring
        t = np.concatenate(([0.0], np.cumsum(10.0 ** rng.uniform(-15, -10, count))))[:count]
        vals = np.ascontiguousarray(np.cumsum(rng.normal(size=(count, nh)), axis=0))
        ring_t = np.zeros(R)
        ring_v = np.zeros((R, nh))
        ring_t[:count], ring_v[:count] = t, vals
        lst = ref.listNew(None)
        for q in range(count):
            rec = ref.historyNew(float(t[q]), vals[q].ctypes.data, nh, 0)
            assert rec and ref.listAdd(lst, rec, None) == 0
        hi = ref.historyInterpNew(None, lst)
        assert hi and ref.historyInterpInitialize(hi) == 0
        hint = C.c_int(0)
        tp = 0.0
        for k in range(80):
            u = rng.random()
            if u < 0.6:                                   # the delayed time creeps forward, as in a run
                tp = tp + 10.0 ** rng.uniform(-14, -10.5)
            elif u < 0.8:
                tp = float(rng.uniform(0.0, t[-1] * 1.1 + 1e-13))
            elif u < 0.9:
                tp = float(t[int(rng.integers(0, count))])
            else:
                tp = -float(rng.uniform(0, 1e-12))        # before the first record
            col = int(rng.integers(0, nh))
            want = C.c_double()
            if tp >= t[-1]:
                # At / past the newest record the reference means to use the last two
                # records (history_interp.c:58-61) but its listNodeGetNext dereferences
                # the missing next node (libs/data/list.c:75-84) and crashes: the
                # intended rule is checked against its formula instead.
                if count < 2:
                    rc0 = -1
                else:
                    tP, tN, xP, xN = t[-2], t[-1], vals[-2, col], vals[-1, col]
                    want.value = ((xN - xP) / (tN - tP)) * (tp - tP) + xP
                    rc0 = 0
                    past += 1
            else:
                ref.ref_shim_quiet(1)
                rc0 = ref.historyInterpSetTime(hi, tp)
                if rc0 == 0:
                    rc0 = ref.historyInterpGetData(hi, col + 1, C.byref(want))
                ref.ref_shim_quiet(0)
            got = C.c_double()
            rc1 = dev.lookup(ring_t.ctypes.data, ring_v.ctypes.data, nh, R, count, C.byref(hint), tp, col,
                             C.byref(got))
            assert (rc0 != 0) == (rc1 != 0), (trial, k, tp, count, rc0, rc1)
            if rc0 == 0:
                assert _bits(want.value) == _bits(got.value), (trial, k, tp, want.value, got.value)
                same += 1
            else:
                failed += 1
                assert ref.historyInterpInitialize(hi) == 0       # the reference forgets its node on error
    assert same > 3000 and failed > 50 and past > 100
