"""The production kernel's translation unit, compiled for the HOST, against the oracle.

Everything a launch executes on the GPU - the device library (simcu_device.cuh),
the code generated for the topology (device records, straight-line LU, expression
code) and the attempt loop itself (simcu_jit_kernel.cuh: claim, operating point,
step / integrate / Newton / LTE / accept-reject, history ring, gridded output) -
is ONE translation unit that NVRTC compiles for sm_100a.  Here the same text is
compiled by g++ behind a 30-line shim for the CUDA built-ins and run as a single
lane over a small batch, on the host-side arrays Prepare would upload
(simcuDebugProgram).  The oracle eliminates with the same pivot sequence
(eoGetPivots -> simcuSetPivots / eoSetPivots); every unknown at every accepted
step of every instance must then be bit-identical, libm calls included (both
sides use glibc here; on the GPU they differ in the last ulp, tests/test_gpu_*).

TEST INFRASTRUCTURE ONLY.  This is not a CPU path of the product - the engine
has none and fails loudly without a CUDA device (tests/test_bench_contract.py);
nothing outside tests/ can reach this harness.  It exists so that the device
code the GPU runs is checked against the reference's algorithm on every CPU
test run, change by change.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import backends
import circuits
import workloads
from eispice_b200 import engine, units

SHIM = r"""
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <climits>
using std::isnan; using std::isfinite;
#define __device__
#define __noinline__
#define __forceinline__ inline
#define __global__
#define __launch_bounds__(a, b)
struct double2 { double x, y; };
struct int4 { int x, y, z, w; };
struct uint3 { unsigned x, y, z; };
static uint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {32, 1, 1}, gridDim = {1, 1, 1};
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline unsigned __ballot_sync(unsigned, int pred) { return pred ? 1u : 0u; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __syncthreads_or(int v) { return v; }
static inline void __syncthreads() {}
static inline int atomicAdd(int *p, int v) { int o = *p; *p += v; return o; }
static inline int min(int a, int b) { return a < b ? a : b; }
"""

DRIVER = r"""
extern "C" int runKernel(int M, int ngrid, int nprobes, int histRecords, int rawMaxPoints,
		const SimcuControl *ctl, const int *pmap, const double *pconst, const double *pinst,
		const int *iinst, const double *tabP, const int *tabDesc, double *Ht, double *Hx,
		double *outGrid, double *outRaw, int *CI, double *outMeasure, int opOnly,
		const int *replayOff, const double *replayTime, const int *replayFlag, const int *replayLinOff,
		const int *replayLin)
{
	SimcuArgs a;
	std::memset(&a, 0, sizeof a);
	a.M = M; a.N = SIMCU_N; a.nnzA = SIMCU_NNZ; a.ctl = *ctl;
	a.pmap = pmap; a.pconst = pconst; a.pinst = pinst; a.iinst = iinst;
	a.tabP = tabP; a.tabDesc = tabDesc;
	a.nh = SIMCU_NH; a.histRecords = histRecords; a.Ht = Ht; a.Hx = Hx;
	a.nprobes = nprobes; a.ngrid = ngrid; a.outGrid = outGrid;
	a.rawMaxPoints = rawMaxPoints; a.rawFirst = 0; a.rawCount = M; a.outRaw = outRaw;
	a.CI = CI; a.outMeasure = outMeasure;
	a.maxAttempts = LLONG_MAX;
	a.replayOff = replayOff; a.replayTime = replayTime; a.replayFlag = replayFlag;
	a.replayLinOff = replayLinOff; a.replayLin = replayLin;
	int counter = 0;
	a.counter = &counter;
	a.lanes = 1; a.count = M; a.claimBelow = 31;       /* one lane claims instance after instance */
	simcuJitKernel(a, opOnly);
	return counter;
}
"""


def _debug_array(eng, what, dtype, tstep=1.0, tstop=1.0):
    L = engine.lib()
    L.simcuDebugProgram.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    data, n = C.c_void_p(), C.c_int()
    assert L.simcuDebugProgram(eng.h, what.encode(), tstep, tstop, 0.0, C.byref(data), C.byref(n)) == 0
    raw = C.string_at(data, n.value) if n.value else b""
    L.simcuFree(data)
    arr = np.frombuffer(raw, dtype=dtype).copy()
    return arr if arr.size else np.zeros(1, dtype=dtype)


def _run_on_host(tmp_path, tag, nl, M, params, probes, tstep, tstop, pivots, op_only=False, options=None,
                 replay=None):
    """Raw records [instance][point][1 + N] of the generated kernel run on the host."""
    eng = engine.Engine(nl, M)
    for k, v in (options or {}).items():
        eng.set_option(k, v)
    if params:
        eng.set_params(params)
    for phase, (pc, pr) in enumerate(pivots):
        eng.set_pivots(phase, pc, pr)
    src, _ = eng.compile_only(probes, source_only=True)
    n = eng.num_unknowns()
    arrays = {k: _debug_array(eng, k, dt) for k, dt in (("pmap", np.int32), ("pconst", np.float64),
              ("pinst", np.float64), ("iinst", np.int32), ("tabP", np.float64), ("tabDesc", np.int32))}
    ctl = _debug_array(eng, "control", np.uint8, tstep, tstop)
    eng.close()
    cpp, so = tmp_path / ("kernel_%s.cpp" % tag), tmp_path / ("kernel_%s.so" % tag)
    cpp.write_text(SHIM + src + DRIVER)
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-w", "-fpermissive", "-shared", "-fPIC", "-o", str(so),
                    str(cpp)], check=True)
    lib = C.CDLL(str(so))
    lib.runKernel.argtypes = [C.c_int] * 5 + [C.c_void_p] * 13 + [C.c_int] + [C.c_void_p] * 5
    nh = int(__import__("re").search(r"#define SIMCU_NH (\d+)", src).group(1))
    ng = int(np.floor(tstop / tstep * (1.0 + 1e-12))) + 1
    if (ng - 1) * tstep < tstop * (1.0 - 1e-9):
        ng += 1
    if op_only:
        ng = 0                                    # as Prepare does for an operating point
    R, raw_max = 1 << 15, 1 << 15
    Ht = np.zeros(R)
    Hx = np.zeros((R, max(nh, 1)))
    out_grid = np.full((max(ng, 1), len(probes), M), np.nan)
    out_raw = np.full((raw_max, n + 1, M), np.nan)
    CI = np.zeros((16, M), dtype=np.int32)
    meas = np.zeros((1, M))
    p = lambda a: a.ctypes.data
    claimed = lib.runKernel(M, ng, len(probes), R, raw_max, p(ctl), p(arrays["pmap"]), p(arrays["pconst"]),
                            p(arrays["pinst"]), p(arrays["iinst"]), p(arrays["tabP"]), p(arrays["tabDesc"]),
                            p(Ht), p(Hx), p(out_grid), p(out_raw), p(CI), p(meas), 1 if op_only else 0,
                            *([p(x) for x in replay] if replay else [None] * 5))
    assert claimed == M + 1                       # M instances + the claim that found the queue empty
    status, npts = CI[1], CI[8]                   # CI_STATUS, CI_NPTS (simcu_program.h)
    if replay:
        return [out_raw[:npts[i], :, i] for i in range(M)], status, CI[13] + CI[15]   # CI_REPLAY_DIFF + CI_LIN_DIFF
    return [out_raw[:npts[i], :, i] for i in range(M)], status, out_grid[:ng]


def _oracle(nl, vi_set, tstep, tstop, pivots=None, op_only=False):
    ora = backends.OracleSim(nl, vi_set)
    if pivots:
        for phase, pv in enumerate(pivots):
            if pv is not None:
                ora.set_pivots(phase, pv[0], pv[1])
    data, names = ora.op() if op_only else ora.tran(tstep, tstop)
    return ora, data


CASES = {
    # name: (workload, instances) or golden circuit
    "cfg1": ("workload", 3), "cfg4": ("workload", 6), "cfg2": ("workload", 2), "cfg3": ("workload", 2),   # cfg4: all six corners
    "cfg5": ("workload", 3),
    "tline_lossy": ("golden", 1), "b_capacitor": ("golden", 1), "w_gauss": ("golden", 1), "w_pwc": ("golden", 1),
    "diode": ("golden", 1), "ibis_ramp": ("golden", 1), "b_voltage_time": ("golden", 1), "realc": ("golden", 1),
    "vi_spline_ta": ("golden", 1), "pwl_current_into_rlc": ("golden", 1),
    # operating points
    "vi_table": ("golden", 1), "ccvs": ("golden", 1), "subckt_nested": ("golden", 1), "b_current": ("golden", 1),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_generated_kernel_on_the_host_equals_the_oracle_bit_for_bit(name, tmp_path):
    kind, M = CASES[name]
    op_only = False
    if kind == "workload":
        w = workloads.make(name, M)
        nl, probes, tstep, tstop = w.cct.netlist(), w.probes, w.tstep, w.tstop
        instances = [w.instance_netlist(i) for i in range(M)]
        params = w.params
    else:
        entry = circuits.all_cases().get(name) or circuits.EXTRA_CASES[name]
        build, an = entry[0], entry[1]
        nl = build().netlist()
        op_only = (an[0] == "op")
        tstep, tstop = (1.0, 1.0) if op_only else (units.float(an[1]), units.float(an[2]))
        instances, params = [(nl, 0)], None
        e0 = engine.Engine(nl, 1)
        probes = e0.variables()[1:3]
        e0.close()
    # the pivot sequences of instance 0 under the reference's partial pivoting
    first, _ = _oracle(instances[0][0], instances[0][1], tstep, tstop, op_only=op_only)
    pivots = [first.first_pivots(0), first.first_pivots(1)]
    assert pivots[0] is not None and (op_only or pivots[1] is not None)
    raw, status, grid = _run_on_host(tmp_path, name, nl, M, params, probes, tstep, tstop,
                                     [pv for pv in pivots if pv is not None], op_only)
    assert (status == 0).all(), status
    points = 0
    for i, (inl, vi_set) in enumerate(instances):
        _, want = _oracle(inl, vi_set, tstep, tstop, pivots, op_only)
        got = raw[i]
        assert got.shape == want.shape, (name, i, got.shape, want.shape)
        assert np.array_equal(got, want), (name, i, float(np.abs(got - want).max()))
        points += want.shape[0]
    assert points >= (1 if op_only else 10)
    if name == "cfg5":
        assert points > 3 * 5000
    assert np.isfinite(grid).all()


@pytest.mark.parametrize("name", ["cfg5", "cfg4", "cfg2"])
@pytest.mark.parametrize("options", [{"solve_blocks": 0, "shared_break_clock": 0}, {"solve_blocks": 1},
                                     {"fmad": 1}, {"growth_limit": 0}],
                         ids=["plain_lu_own_clocks", "blocks_every_solve", "fmad_option", "no_growth_guard"])
def test_kernel_options_do_not_change_a_bit_on_the_host(name, options, tmp_path):
    """The all-at-once LU with per-device break-point clocks, the block-ordered LU
    without the once-per-loop shortcut, the build without the growth guard - and,
    on a host without contraction, the fmad option - walk the same numbers as the
    default kernel: equal to the oracle bit for bit."""
    M = 2
    w = workloads.make(name, M)
    instances = [w.instance_netlist(i) for i in range(M)]
    first, _ = _oracle(instances[0][0], instances[0][1], w.tstep, w.tstop)
    pivots = [first.first_pivots(0), first.first_pivots(1)]
    raw, status, _ = _run_on_host(tmp_path, name, w.cct.netlist(), M, w.params, w.probes, w.tstep, w.tstop,
                                  pivots, options=options)
    assert (status == 0).all()
    for i, (inl, vi_set) in enumerate(instances):
        _, want = _oracle(inl, vi_set, w.tstep, w.tstop, pivots)
        assert np.array_equal(raw[i], want), (name, options, i)


@pytest.mark.parametrize("name", ["cfg5", "cfg4", "cfg2"])
def test_replay_kernel_on_the_host_follows_the_oracle_log_without_a_disagreement(name, tmp_path):
    """The kernel compiled with option replay takes every accept / reject decision
    and every convergence-test outcome from the oracle's log (what
    tests/test_gpu_replay.py does on the GPU) - and, walking the same numbers,
    would have decided the same each time: zero disagreements, same records."""
    M = 2
    w = workloads.make(name, M)
    instances = [w.instance_netlist(i) for i in range(M)]
    first, _ = _oracle(instances[0][0], instances[0][1], w.tstep, w.tstop)
    pivots = [first.first_pivots(0), first.first_pivots(1)]
    logs, want = [], []
    for inl, vi_set in instances:
        ora = backends.OracleSim(inl, vi_set)
        for phase, pv in enumerate(pivots):
            ora.set_pivots(phase, pv[0], pv[1])
        ora.log_attempts()
        data, _ = ora.tran(w.tstep, w.tstop)
        t, f = ora.attempt_log()
        logs.append((t, f, ora.linearize_log()))
        want.append(data)
    off = np.zeros(M + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(x[0]) for x in logs])
    loff = np.zeros(M + 1, dtype=np.int32)
    loff[1:] = np.cumsum([len(x[2]) for x in logs])
    replay = [off, np.ascontiguousarray(np.concatenate([x[0] for x in logs])),
              np.ascontiguousarray(np.concatenate([x[1] for x in logs]).astype(np.int32)), loff,
              np.ascontiguousarray(np.concatenate([x[2] for x in logs]).astype(np.int32))]
    raw, status, diffs = _run_on_host(tmp_path, name, w.cct.netlist(), M, w.params, w.probes, w.tstep, w.tstop,
                                      pivots, options={"replay": 1}, replay=replay)
    assert (status == 0).all() and (diffs == 0).all(), (status, diffs)
    for i in range(M):
        assert np.array_equal(raw[i], want[i]), (name, i)
