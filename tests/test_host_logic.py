"""Host side of libsimulator_cuda without a GPU: the library loads and exports
every symbol of include/simulator_cuda.h, the netlist builder creates the
reference's rows and sparsity pattern (checked against the oracle, which is
pinned to the reference), the expression compiler + interpreter agree with the
reference's calculon (golden vectors + fuzz against the oracle), and the
symbolic LU analysis yields usable pivot sequences."""
import json
import math
import os
import random
import re

import numpy as np
import pytest

import backends
import circuits
from eispice_b200 import engine, units

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    lib = engine.lib()
    hdr = open(os.path.join(ROOT, "include", "simulator_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(simcu[A-Z]\w*)\s*\(", hdr))
    assert len(names) >= 30
    for n in sorted(names):
        assert hasattr(lib, n), n


@pytest.mark.parametrize("name", sorted(circuits.all_cases()))
def test_rows_and_pattern_match_reference_order(name):
    build = circuits.all_cases()[name][0]
    nl = build().netlist()
    eng = engine.Engine(nl, 1)
    ora = backends.OracleSim(nl)
    n, ri, cs = eng.pattern()
    n2, ri2, cs2 = ora.pattern()
    assert n == n2
    assert np.array_equal(cs, cs2) and np.array_equal(ri, ri2)
    # names: run-free comparison through a tiny OP on the oracle
    _, names = ora.op() if True else (None, None)
    assert eng.variables() == names
    eng.close()


def test_multiple_table_sets_keep_the_pattern():
    cct = circuits._new("sets")
    import eispice_b200 as eispice
    t1 = eispice.PWL([[-1, -1], [0, 0], [1, 2]])
    t2 = eispice.PWL([[-1, -2], [0, 0], [1, 3]])
    cct.Vx = eispice.V(1, 0, 1)
    cct.VIx = eispice.VI(1, 0, [t1, t2])
    eng = engine.Engine(cct.netlist(), 4)
    eng.set_params({"VIx.set": np.array([0, 1, 1, 0])})
    with pytest.raises(KeyError):
        eng.set_params({"VIx.bogus": np.zeros(4)})
    assert eng.num_unknowns() == 4
    eng.close()


def test_rejects_what_the_reference_rejects():
    L = engine.lib()
    import ctypes as C
    L.simcuSetOption(None, b"quiet", 1.0)
    try:
        h = C.c_void_p(L.simcuNew(None, 1))
        one = C.c_double(1.0)
        assert L.simcuAddResistor(h, b"R1", b"a", b"0", C.byref(one)) == 0
        assert L.simcuAddResistor(h, b"R1", b"a", b"b", C.byref(one)) == -1      # listed twice
        assert L.simcuAddSource(h, b"V1", b"0", b"0", b"v", C.byref(one), b"\0", None) == -1
        assert L.simcuAddSource(h, b"V2", b"a", b"0", b"x", C.byref(one), b"\0", None) == -1
        assert L.simcuAddTLine(h, b"T1", b"0", b"0", b"b", b"0", C.byref(one), C.byref(one),
                               C.byref(one)) == -1
        eq = C.create_string_buffer(b"foo + 1")
        assert L.simcuAddNonlinearSource(h, b"B1", b"a", b"0", b"i", eq) == -1   # bare variable
        eq2 = C.create_string_buffer(b"v(a) * ")
        assert L.simcuAddNonlinearSource(h, b"B2", b"a", b"0", b"i", eq2) == -1  # syntax
        assert L.simcuAddCallbackSource(h, b"P1", b"a", b"0", b"i", None, None, None, 0,
                                        None, None) == -1
        assert L.simcuNew(h, 1) is None                                          # arg must be NULL
        L.simcuDestroy(C.byref(h))
    finally:
        L.simcuSetOption(None, b"quiet", 0.0)


# ---- expression compiler ---------------------------------------------------
VEC = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "calc_vectors.json")))


def _same(a, b, rel=1e-12):
    if a is None or b is None:
        return a is None and b is None
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    if math.isinf(a) or math.isinf(b):
        return a == b
    return abs(a - b) <= rel * max(abs(a), abs(b)) + 1e-300


def test_bytecode_matches_reference_golden_vectors():
    names = VEC["names"]
    for case in VEC["cases"]:
        rc, val, der = engine.calc_eval(case["expr"], names, case["values"])
        assert (rc == 0) == case["ok"], case["expr"]
        if rc == 0:
            assert _same(val, case["value"]), (case["expr"], val, case["value"])
            for d, g in zip(der, case["derivs"]):
                assert _same(None if math.isnan(d) else d, g), (case["expr"], der, case["derivs"])


def _random_expr(rng, depth=0):
    atoms = ["v(a)", "v(b)", "i(Vx)", "time", "2", "0.5", "1m", "3k", "1e-3", "v(a,b)",
             "2.5Meg", "4mil", "7u"]
    funcs = ["abs", "sin", "cos", "exp", "sqrt", "uramp", "u", "atan", "sinh", "ln", "log",
             "tan", "cosh", "asinh"]
    if depth > 3 or rng.random() < 0.25:
        return rng.choice(atoms)
    r = rng.random()
    if r < 0.45:
        return "%s %s %s" % (_random_expr(rng, depth + 1), rng.choice("+-*/^"),
                             _random_expr(rng, depth + 1))
    if r < 0.6:
        return "%s(%s)" % (rng.choice(funcs), _random_expr(rng, depth + 1))
    if r < 0.7:
        return "(%s)" % _random_expr(rng, depth + 1)
    if r < 0.78:
        return "-%s" % _random_expr(rng, depth + 1)
    if r < 0.86:
        return "%s(%s)" % (_random_expr(rng, depth + 1), _random_expr(rng, depth + 1))
    if r < 0.93:
        return "if(%s %s %s)" % (_random_expr(rng, depth + 1),
                                 rng.choice([">", "<", ">=", "<=", "==", "!="]),
                                 _random_expr(rng, depth + 1))
    return "if((%s > %s) %s (%s < %s))" % (
        _random_expr(rng, depth + 1), _random_expr(rng, depth + 1), rng.choice(["&&", "||"]),
        _random_expr(rng, depth + 1), _random_expr(rng, depth + 1))


def test_bytecode_fuzz_against_oracle():
    rng = random.Random(20261)
    names = ["v(a)", "v(b)", "i(Vx)", "time"]
    agree = 0
    for _ in range(1500):
        expr = _random_expr(rng)
        values = [rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-1, 1), rng.uniform(0, 1e-6)]
        rc_o, v_o, d_o = backends.eval_expr("oracle", expr, names, values)
        rc_e, v_e, d_e = engine.calc_eval(expr, names, values)
        assert (rc_o == 0) == (rc_e == 0), (expr, values, rc_o, rc_e)
        if rc_o == 0:
            assert _same(v_o, v_e), (expr, values, v_o, v_e)
            for a, b in zip(d_o, d_e):
                assert _same(a, b), (expr, values, d_o, d_e)
            agree += 1
    assert agree > 300


def test_expression_limits():
    deep = "(" * 30 + "v(a)" + ")" * 30                     # nesting alone costs no stack
    assert engine.calc_eval(deep, ["v(a)"], [1.5])[0] == 0
    chain = "+".join(["(v(a)*" * 30 + "1" + ")" * 30])        # 31 live operands
    assert engine.calc_eval(chain, ["v(a)"], [1.0])[0] == -1


# ---- symbolic analysis -----------------------------------------------------
def _dense_solve_with_pivots(A, b, pc, pr):
    """Replay a fixed (column, row) sequence densely -- what the kernel does."""
    W = A.copy()
    y = b.copy()
    n = len(b)
    rdone = np.zeros(n, bool)
    cdone = np.zeros(n, bool)
    for k in range(n):
        c, p = pc[k], pr[k]
        rdone[p] = True
        cdone[c] = True
        for r in range(n):
            if rdone[r]:
                continue
            l = W[r, c] / W[p, c]
            W[r, ~cdone] -= l * W[p, ~cdone]
            y[r] -= l * y[p]
    x = np.zeros(n)
    for k in range(n - 1, -1, -1):
        c, p = pc[k], pr[k]
        acc = y[p]
        for j in range(k + 1, n):
            acc -= W[p, pc[j]] * x[pc[j]]
        x[c] = acc / W[p, c]
    return x


@pytest.mark.parametrize("seed", range(6))
def test_symbolic_analysis_on_mna_like_matrices(seed):
    import ctypes as C
    rng = np.random.default_rng(seed)
    n = int(rng.integers(5, 40))
    # conductance-like block with zero-diagonal branch rows (voltage sources)
    S = np.zeros((n, n), bool)
    for i in range(n):
        S[i, i] = rng.random() < 0.8
    for _ in range(3 * n):
        i, j = rng.integers(0, n, 2)
        S[i, j] = S[j, i] = True
    nb = max(1, n // 5)
    for k in range(nb):            # branch k couples to node k with +-1, zero diagonal
        b = n - 1 - k
        S[b, :] = False
        S[:, b] = False
        S[b, k] = S[k, b] = True
    pilots = []
    for q in range(3):
        A = np.where(S, rng.uniform(0.5, 2.0, (n, n)) * 10.0 ** rng.integers(-3, 3, (n, n)), 0.0)
        for i in range(n - nb):
            A[i, i] = np.abs(A[i]).sum() + 1.0
            S[i, i] = True
        for k in range(nb):
            b = n - 1 - k
            A[b, k] = 1.0
            A[k, b] = 1.0
        pilots.append(A)
    cs = [0]
    ri = []
    for c in range(n):
        rows = np.nonzero(S[:, c])[0]
        ri += list(rows)
        cs.append(len(ri))
    cs = np.array(cs, dtype=np.int32)
    ri = np.array(ri, dtype=np.int32)
    vals = np.array([[A[r, c] for c in range(n) for r in np.nonzero(S[:, c])[0]] for A in pilots])
    pc = np.zeros(n, dtype=np.int32)
    pr = np.zeros(n, dtype=np.int32)
    slots, nf = C.c_int(), C.c_int()
    rc = engine.lib().simcuSymbolic(n, cs.ctypes.data, ri.ctypes.data, len(pilots),
                                    np.ascontiguousarray(vals).ctypes.data, 0.1,
                                    pc.ctypes.data, pr.ctypes.data, C.byref(slots), C.byref(nf))
    assert rc == 0
    assert sorted(pc) == list(range(n)) and sorted(pr) == list(range(n))
    assert slots.value >= len(ri) and nf.value >= n
    for A in pilots:
        b = rng.normal(size=n)
        x = _dense_solve_with_pivots(A, b, pc, pr)
        assert np.allclose(A @ x, b, rtol=1e-8, atol=1e-8 * np.abs(b).max())


def test_symbolic_analysis_reports_singular():
    import ctypes as C
    L = engine.lib()
    L.simcuSetOption(None, b"quiet", 1.0)
    try:
        cs = np.array([0, 1, 2], dtype=np.int32)
        ri = np.array([0, 0], dtype=np.int32)       # both entries in row 0
        vals = np.array([1.0, 2.0])
        pc = np.zeros(2, dtype=np.int32)
        pr = np.zeros(2, dtype=np.int32)
        rc = L.simcuSymbolic(2, cs.ctypes.data, ri.ctypes.data, 1, vals.ctypes.data, 0.1,
                             pc.ctypes.data, pr.ctypes.data, None, None)
        assert rc == -1
    finally:
        L.simcuSetOption(None, b"quiet", 0.0)


def test_runs_fail_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = engine.lib()
    L.simcuSetOption(None, b"quiet", 1.0)
    try:
        cct = circuits.cfg1_rc()
        with pytest.raises(RuntimeError):
            cct.tran('0.01n', '10n')
        with pytest.raises(RuntimeError):
            cct.tran_batch('0.01n', '10n', probes=["v(out)"])
    finally:
        L.simcuSetOption(None, b"quiet", 0.0)


# ---- topology-specialised kernel: generation + NVRTC, no GPU needed ---------
@pytest.mark.parametrize("name", ["cfg1_rc", "cfg2_rectifier", "cfg3_oscillator", "tline_lossy",
                                  "b_capacitor", "b_voltage_time", "w_pwc", "w_gauss", "realc",
                                  "vi_table", "cccs", "cfg4_ibis"])
def test_specialised_kernel_compiles_for_sm100a(name):
    build, an, _ = circuits.all_cases()[name]
    eng = engine.Engine(build().netlist(), 1)
    names = eng.variables()
    src, cubin_bytes = eng.compile_only(names[1:3])
    assert cubin_bytes > 10000
    assert "simcuJitKernel" in src and "genFactorSolve1" in src
    # every device of the netlist appears in the load phase, in insertion order
    load = src[src.index("DEV void genLoad"):src.index("DEV void genInitStep")]
    assert load.count("devLoad(c, d, bv)") == len(build().netlist())
    eng.close()


def test_specialised_source_is_independent_of_instances_and_values():
    """One compilation serves every batch size and parameter value."""
    import workloads
    w1 = workloads.make("cfg2", 16)
    w2 = workloads.make("cfg2", 48, seed_offset=3)
    srcs = []
    for w in (w1, w2):
        eng = engine.Engine(w.cct.netlist(), w.M)
        eng.set_params(w.params)
        srcs.append(eng.compile_only(w.probes)[0])
        eng.close()
    assert srcs[0] == srcs[1]


# ---- generated LU: compile the emitted straight-line code for the HOST and
# ---- check it against a dense solve ------------------------------------------
def _host_lu_harness(src, tmp_path, phase, tag=""):
    """Builds a shared object around genFactorSolve<phase> of a generated source.
    solve(A, B, X, first): X is in/out, so a later solve of a Newton loop
    (first = 0) can leave the blocks no linearize stamp reaches as they are."""
    import ctypes as C
    import subprocess
    name = "genFactorSolve%d" % phase
    body = src[src.index("DEV void %s(Inst &c, const bool first)" % name):]
    body = body[:body.index("\n}\n") + 3]
    n = int(re.search(r"#define SIMCU_N (\d+)", src).group(1))
    nnz = int(re.search(r"#define SIMCU_NNZ (\d+)", src).group(1))
    code = """
#include <cmath>
#include <cstring>
#define DEV static inline
#define SIMCU_ERR_SINGULAR 2
#define SIMCU_LMAX 1e6
using std::isfinite;
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
struct Inst { double Av[%d], Bv[%d], Xv[%d]; int err, growth; bool growInv; };
%s
extern "C" int solve(const double *A, const double *B, double *X, int first)
{
	Inst c; c.err = 0; c.growth = 0; c.growInv = false;
	std::memcpy(c.Av, A, sizeof c.Av); std::memcpy(c.Bv, B, sizeof c.Bv);
	std::memcpy(c.Xv, X, sizeof c.Xv);
	%s(c, first != 0);
	std::memcpy(X, c.Xv, sizeof c.Xv);
	return c.err;
}
""" % (max(nnz, 1), n, n, body, name)
    cpp = tmp_path / ("lu%d%s.cpp" % (phase, tag))
    so = tmp_path / ("lu%d%s.so" % (phase, tag))
    cpp.write_text(code)
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(cpp)],
                   check=True)
    lib = C.CDLL(str(so))
    literals = {int(k): v for k, v in re.findall(r"double w(\d+) = __longlong_as_double\((0x[0-9a-f]+)LL\);", body)}
    loaded = {int(k) for k in re.findall(r"double w(\d+) = c\.Av\[\d+\];", body)}
    return lib, n, nnz, literals, loaded


@pytest.mark.parametrize("name", ["cfg4_ibis", "tline_lossy", "cfg2_rectifier", "realc"])
def test_generated_lu_solves_the_system(name, tmp_path):
    import ctypes as C
    import struct
    build = circuits.all_cases()[name][0]
    eng = engine.Engine(build().netlist(), 1)
    src, _ = eng.compile_only(eng.variables()[1:2])
    n, ri, cs = eng.pattern()
    rng = np.random.default_rng(1)
    for phase in (0, 1):
        lib, n2, nnz, literals, loaded = _host_lu_harness(src, tmp_path, phase)
        assert n2 == n and nnz == len(ri)
        A = np.zeros(nnz)
        for s in loaded:
            A[s] = rng.uniform(0.5, 2.0) * rng.choice([-1.0, 1.0])
        dense = np.zeros((n, n))
        col = np.repeat(np.arange(n), np.diff(cs))
        for s in range(nnz):
            v = A[s]
            if s in literals:
                v = struct.unpack("<d", struct.pack("<Q", int(literals[s], 16)))[0]
            dense[ri[s], col[s]] = v
        # make the structurally chosen pivots numerically safe: diagonal boost on loaded slots
        for s in loaded:
            if ri[s] == col[s]:
                A[s] = dense[ri[s], col[s]] = 10.0 + abs(A[s])
        if abs(np.linalg.det(dense)) < 1e-9:
            continue
        b = rng.normal(size=n)
        x = np.zeros(n)
        err = lib.solve(A.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                        x.ctypes.data_as(C.c_void_p), 1)
        assert err == 0
        want = np.linalg.solve(dense, b)
        assert np.allclose(x, want, rtol=1e-8, atol=1e-9 * np.abs(want).max()), (name, phase)
    eng.close()


GOLDEN_WITH_BLOCKS = ["b_capacitor", "b_current", "b_voltage_time", "cccs", "ccvs", "diode", "ibis_ramp",
                      "subckt_nested", "tline", "tline_lossy", "vccs", "vcvs", "vi_table", "vi_table_neg",
                      "cfg3_oscillator", "cfg5_ibis_lossy"]


@pytest.mark.parametrize("name", ["cfg5", "cfg4", "cfg2", "cfg5_masked"] + GOLDEN_WITH_BLOCKS)
def test_block_ordered_lu_is_bit_identical_and_later_solves_skip_only_invariant_blocks(name, tmp_path):
    """The generated LU regrouped block by block (solve_blocks = 1) and with the
    blocks no linearize stamp reaches run only in the first solve of a Newton
    loop (2, the default) gives, bit for bit, the solution of the all-at-once
    emission (0) - also in a LATER solve, after everything a linearize pass can
    change (the slots and rows the generated code lists) has changed."""
    import ctypes as C
    import workloads
    srcs = {}
    for blocks in (0, 1, 2):
        if name in GOLDEN_WITH_BLOCKS:         # a golden circuit of the reference's tests
            eng = engine.Engine(circuits.all_cases()[name][0]().netlist(), 1)
            probes = eng.variables()[1:2]
        else:                                  # a BASELINE sweep
            w = workloads.cfg5_masked(16) if name == "cfg5_masked" else workloads.make(name, 16)
            eng = engine.Engine(w.cct.netlist(), w.M)
            eng.set_params(w.params)
            probes = w.probes
        eng.set_option("solve_blocks", blocks)
        srcs[blocks] = eng.compile_only(probes, source_only=True)[0]
        eng.close()
    rng = np.random.default_rng(11)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    for phase in (0, 1):
        libs = {b: _host_lu_harness(srcs[b], tmp_path, phase, "b%d" % b) for b in (0, 1, 2)}
        _, n, nnz, _, loaded = libs[0]
        body = srcs[2][srcs[2].index("/* %s:" % ("transient" if phase else "operating point")):]
        m = re.search(r"/\* linearize can change A slots:([ \d]*); B rows:([ \d]*) \*/", body)
        tslots = [int(v) for v in m.group(1).split()]
        trows = [int(v) for v in m.group(2).split()]
        ninv = int(re.search(r"(\d+) of them not reached by any linearize stamp", body).group(1))
        if name == "cfg5" and phase == 1:
            assert ninv >= 7            # the inner nodes of the 8-section line
        if name == "cfg2":
            assert ninv == 0 or len(tslots) > 0
        for trial in range(12):
            A = np.zeros(max(nnz, 1))
            for sl in loaded:
                A[sl] = rng.uniform(0.5, 2.0) * rng.choice([-1.0, 1.0])
            b = rng.normal(size=n)
            xs = {}
            for blk, (lib, *_rest) in libs.items():
                x = np.full(n, np.nan)
                lib.solve(ptr(A), ptr(b), ptr(x), 1)
                xs[blk] = x
            assert xs[0].tobytes() == xs[1].tobytes() == xs[2].tobytes(), (name, phase, trial)
            # a later solve of the same loop: only what linearize can change changed
            A2, b2 = A.copy(), b.copy()
            for sl in tslots:
                if sl in loaded:
                    A2[sl] = rng.uniform(0.5, 2.0) * rng.choice([-1.0, 1.0])
            for r in trows:
                b2[r - 1] = rng.normal()
            full = np.full(n, np.nan)
            libs[0][0].solve(ptr(A2), ptr(b2), ptr(full), 1)
            later = xs[2].copy()
            libs[2][0].solve(ptr(A2), ptr(b2), ptr(later), 0)
            assert full.tobytes() == later.tobytes(), (name, phase, trial)


# ---- generated expression code: compile genCalc / genCalcGrad for the HOST and
# ---- compare with the bytecode interpreter (itself pinned to the reference) ----
def _host_calc_harness(src, tmp_path, n_unknowns):
    import ctypes as C
    import subprocess
    gmin = re.search(r"#define SIMCU_GMIN \((.*)\)", src).group(1)
    start = src.index("DEV void genCalc(Inst &c, int id, int dvar, double &f, double &dv)\n{")
    end = src.index("DEV void genLoad(Inst &c)")
    body = src[start:end]
    calc_h = open(os.path.join(ROOT, "eispice_b200", "csrc", "simcu_calc.h")).read()
    prog_h = open(os.path.join(ROOT, "eispice_b200", "csrc", "simcu_program.h")).read()
    calc_h = "\n".join(ln for ln in calc_h.splitlines() if not ln.lstrip().startswith("#include"))
    code = """
#include <cmath>
#include <cstring>
%s
%s
#define DEV static inline
#define SIMCU_GMIN (%s)
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
struct Inst { double x[%d]; double time; double X(int row) const { return row ? x[row] : 0.0; } };
%s
extern "C" void eval(int id, int nvars, const double *x, double time, double *f, double *g, double *g1)
{
	Inst c; std::memcpy(c.x, x, sizeof c.x); c.time = time;
	double dv;
	genCalc(c, id, -1, *f, dv);
	genCalcGrad(c, id, g);
	for (int v = 0; v < nvars; v++) { double ff; genCalc(c, id, v, ff, g1[v]); }
}
""" % (prog_h, calc_h, gmin, n_unknowns + 1, body)
    cpp = tmp_path / "calc.cpp"
    so = tmp_path / "calc.so"
    cpp.write_text(code)
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(cpp)],
                   check=True)
    return C.CDLL(str(so))


@pytest.mark.parametrize("name", ["cfg2_rectifier", "b_current", "b_capacitor", "cfg3_oscillator",
                                  "b_voltage_time", "cccs"])
def test_generated_expression_code_equals_the_interpreter(name, tmp_path):
    import ctypes as C
    build = circuits.all_cases()[name][0]
    nl = build().netlist()
    eng = engine.Engine(nl, 1)
    names = eng.variables()                    # names[row], row 0 = "time" stands for ground here
    src, _ = eng.compile_only(names[1:2])
    lib = _host_calc_harness(src, tmp_path, len(names) - 1)
    exprs = [d[5] for d in nl if d[0] == "B"]
    cases = re.findall(r"\tcase (\d+): \{\n((?:\t\tconst double v\d+ = [^\n]+\n)*)\t\tswitch \(dvar\)", src)
    assert len(cases) == len(exprs)
    rng = np.random.default_rng(4)
    for (cid, decl), expr in zip(cases, exprs):
        refs = re.findall(r"const double v(\d+) = (c\.time|c\.X\((\d+)\));", decl)
        var_names = ["time" if r[1] == "c.time" else names[int(r[2])] for r in refs]
        for _ in range(25):
            x = np.zeros(len(names))
            x[1:] = rng.uniform(-2.0, 2.0, len(names) - 1)
            t = float(rng.uniform(0, 1e-8))
            vals = [t if v == "time" else x[names.index(v)] for v in var_names]
            rc, want_f, want_d = engine.calc_eval(expr, var_names, vals)
            f = C.c_double()
            g = (C.c_double * 16)(*([float("nan")] * 16))
            g1 = (C.c_double * 16)()
            lib.eval(int(cid), len(var_names), x.ctypes.data_as(C.c_void_p), C.c_double(t),
                     C.byref(f), g, g1)
            if rc != 0:
                continue                       # NaN somewhere: both sides flag it
            assert _same(f.value, want_f), (expr, vals, f.value, want_f)
            for k, v in enumerate(var_names):
                if v == "time":
                    continue
                assert _same(g[k], want_d[k]), (expr, v, g[k], want_d[k])
                assert _same(g1[k], want_d[k]), (expr, v, g1[k], want_d[k])
    eng.close()


def test_b_source_variables_create_rows_in_order_of_appearance():
    """A B equation may mention unknowns that no device has created yet: rows
    appear in order of first mention inside the equation
    (devices/nonlinear_i.c:147-192, core/row.c:163-204), and the Jacobian
    entries join the pattern."""
    import eispice_b200 as eispice
    cct = circuits._new("order")
    cct.Bx = eispice.B('out', 0, eispice.Current, "v(later) * 2 + i(Vz) + v(out)")
    cct.Rz = eispice.R('later', 0, 10)
    cct.Vz = eispice.V('later', 0, 1)
    cct.Ro = eispice.R('out', 0, 5)
    nl = cct.netlist()
    eng = engine.Engine(nl, 1)
    assert eng.variables() == backends.OracleSim(nl).op()[1]
    assert eng.variables()[:6] == ["time", "v(out)", "i(Bx)", "v(Bx)", "v(later)", "i(Vz)"]
    n, ri, cs = eng.pattern()
    n2, ri2, cs2 = backends.OracleSim(nl).pattern()
    assert np.array_equal(ri, ri2) and np.array_equal(cs, cs2)
    eng.close()


def test_measure_and_probe_names_are_checked_on_the_host():
    import ctypes as C
    L = engine.lib()
    L.simcuSetOption(None, b"quiet", 1.0)
    try:
        eng = engine.Engine(circuits.cfg1_rc().netlist(), 4)
        with pytest.raises(KeyError):
            eng.set_measures([("v(nowhere)", "max")])
        with pytest.raises(KeyError):
            eng.set_measures([("v(out)", "median")])
        eng.set_measures([("v(out)", "max"), ("i(V1)", "tcross_fall", -1e-3)])
        src, _ = eng.compile_only(["v(out)"])
        assert "#define SIMCU_NMEASURES 2" in src and src.count("measureStep(c, ") == 2
        eng.close()
    finally:
        L.simcuSetOption(None, b"quiet", 0.0)


def test_windows_and_source_table_sets_compile_without_a_gpu():
    """Per-instance analysis windows and PWL-source table sets (batched K-table
    generation): the specialised kernel is generated with SIMCU_WINDOWS and
    compiles for sm_100a; bad input is rejected at the C-ABI."""
    import eispice_b200 as eispice
    cct = circuits._new("pwl sets")
    t0 = [[0, 0], [1e-9, 1], [2e-9, 0]]
    t1 = [[0, 0], [2e-9, 2], [4e-9, 0]]
    cct.V1 = eispice.V('in', 0, 0, eispice.PWL(t0))
    cct.R1 = eispice.R('in', 'out', 50)
    cct.C1 = eispice.C('out', 0, '1p')
    nl = [tuple(d[:6]) + ((d[6][0], d[6][1], [t1]),) if d[1] == "V1" else d for d in cct.netlist()]
    eng = engine.Engine(nl, 3)
    eng.set_params({"V1.set": np.array([0, 1, 1], dtype=np.int32)})
    eng.set_params({"V1.set": np.array([0, 2, 1], dtype=np.int32)})          # only sets 0 and 1 exist
    engine.lib().simcuSetOption(None, b"quiet", 1.0)
    with pytest.raises(RuntimeError):
        eng.compile_only(["v(out)"])
    engine.lib().simcuSetOption(None, b"quiet", 0.0)
    eng.set_params({"V1.set": np.array([0, 1, 1], dtype=np.int32)})
    eng.set_windows([1e-11, 2e-11, 2e-11], [2e-9, 4e-9, 4e-9])
    src, nbytes = eng.compile_only(["v(out)"])
    assert "#define SIMCU_WINDOWS 1" in src and nbytes > 10000
    assert "DF_WSETSLOT" in src
    with pytest.raises(ValueError):
        eng.set_windows([1e-11, -1.0, 1e-11], [2e-9, 4e-9, 4e-9])
    eng.set_windows(None)
    src, _ = eng.compile_only(["v(out)"])
    assert "#define SIMCU_WINDOWS 0" in src
    eng.close()


def test_cubin_cache_on_disk(tmp_path, monkeypatch):
    """SIMCU_CACHE_DIR: the second compile of a topology (any process) reads the
    cubin back instead of running NVRTC again."""
    import time
    monkeypatch.setenv("SIMCU_CACHE_DIR", str(tmp_path))
    w_nl = circuits.all_cases()["cfg4_ibis"][0]().netlist()
    t = []
    sizes = []
    for _ in range(2):
        eng = engine.Engine(w_nl, 4)
        t0 = time.time()
        src, n = eng.compile_only(["v(vo)"])
        t.append(time.time() - t0)
        sizes.append(n)
        eng.close()
    files = [f for f in os.listdir(tmp_path) if f.endswith(".cubin")]
    assert len(files) == 1 and sizes[0] == sizes[1] == os.path.getsize(os.path.join(tmp_path, files[0]))
    assert t[1] < 0.5 * t[0], t


# ---- shared break-point clock: compile the device functions for the HOST and
# ---- compare them with the per-device clocks on random attempt sequences ------
def test_shared_break_clock_equals_the_per_device_clocks(tmp_path):
    """checkbreak.c:43-67 keeps t[] per device; every test that runs at every
    attempt sees the same times, so the specialised kernel keeps ONE clock per
    instance (cbBegin / cbIsBreakClocked).  Same decisions, same state, on random
    sequences of attempts incl. rejected ones (time steps back) and flat signals."""
    import ctypes as C
    import subprocess
    dev = open(os.path.join(ROOT, "eispice_b200", "csrc", "simcu_device.cuh")).read()
    start = dev.index("DEVFN int cbAngleBreak(")
    end = dev.index("#else\n#define cbIsBreakClocked cbIsBreak\nDEV void cbInitialize")
    code = """
#include <cmath>
#define DEV static inline
#define DEVFN static
#define SIMCU_JIT 1
#define SIMCU_SHAREDCB 1
#define HUGE_VAL (__builtin_huge_val())
static double simcuAtan2(double y, double x) { return atan2(y, x); }
struct Inst {
	mutable double sv[32]; mutable int iv[4]; double time;
	mutable double cbT0, cbT1, cbDt1; mutable int cbN; mutable bool cbAdv;
	double &S(int i) const { return sv[i]; }
	int &IS(int i) const { return iv[i]; }
};
%s
extern "C" int run(int n, const double *t, const double *x, double maxAngle, int *own, int *clocked, double *state)
{
	Inst c;
	cbInitialize(c, 0, 0, 0.0); cbInitialize(c, 9, 1, 0.0);
	c.cbT0 = 0.0; c.cbT1 = 0.0; c.cbDt1 = 0.0; c.cbN = 0; c.cbAdv = false;
	for (int k = 0; k < n; k++) {
		c.time = t[k];
		cbBegin(c);
		own[k] = cbIsBreak(c, 0, 0, x[k], maxAngle);
		clocked[k] = cbIsBreakClocked(c, 9, 1, x[k], maxAngle);
		if (c.S(0) != c.S(9) || c.S(1) != c.S(10) || c.S(7) != c.S(16)) return k + 1;
		if (c.S(3) != c.cbT0 || c.S(4) != c.cbT1 || c.S(5) != c.cbDt1 || c.IS(0) != c.cbN) return -(k + 1);
	}
	(void)state;
	return 0;
}
""" % dev[start:end]
    cpp, so = tmp_path / "cb.cpp", tmp_path / "cb.so"
    cpp.write_text(code)
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(cpp)],
                   check=True)
    lib = C.CDLL(str(so))
    lib.run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    breaks = 0
    for trial in range(200):
        n = 400
        t = np.zeros(n)
        now, last_ok = 0.0, 0.0
        for k in range(n):
            u = rng.random()
            if u < 0.2 and k:            # rejected attempt: retry from the last accepted time
                now = last_ok + (now - last_ok) * rng.choice([0.125, 0.5, 0.9])
            else:
                last_ok = now
                now = now + 1e-11 * 10 ** rng.uniform(-3, 1)
            t[k] = now
        kind = trial % 4
        if kind == 0:
            x = np.cumsum(rng.normal(size=n)) * 1e-3
        elif kind == 1:
            x = np.where(rng.random(n) < 0.7, 1.25, rng.normal(size=n))      # mostly flat
        elif kind == 2:
            x = 1e-12 * np.cumsum(rng.normal(size=n))                       # angles near the threshold
        else:
            x = np.sin(2e9 * t) * 3.3
        own = np.zeros(n, dtype=np.int32)
        clk = np.zeros(n, dtype=np.int32)
        rc = lib.run(n, t.ctypes.data, x.ctypes.data, 1.0471975511965976, own.ctypes.data,
                     clk.ctypes.data, None)
        assert rc == 0, (trial, rc)
        assert (own == clk).all(), trial
        breaks += int(own.sum())
    assert breaks > 100


# ---- "only linearize stamps change inside a Newton loop": the claim behind
# ---- solve_blocks = 2, checked against the oracle's stamps on every circuit ----
def _declared_newton_stamps(src, phase):
    body = src[src.index("/* %s:" % ("transient" if phase else "operating point")):]
    m = re.search(r"/\* linearize can change A slots:([ \d]*); B rows:([ \d]*) \*/", body)
    return [int(v) for v in m.group(1).split()], [int(v) for v in m.group(2).split()]


def _check_newton_stamps(nl, vi_set, an, label):
    eng = engine.Engine(nl, 1)
    src, _ = eng.compile_only(eng.variables()[1:2], source_only=True)
    n, ri, cs = eng.pattern()
    col = np.repeat(np.arange(n), np.diff(cs))
    eng.close()
    ora = backends.OracleSim(nl, vi_set)
    ora.watch_stamps()
    if an[0] == "op":
        ora.op()
    else:
        try:
            ora.tran(*[units.float(x) for x in an[1:]])
        except RuntimeError:
            pass                      # an aborted run still watched every loop it made
    seen = 0
    for phase in (0, 1):
        got = ora.stamp_changes(phase)
        if got is None:
            continue
        A, B = got
        slots, rows = _declared_newton_stamps(src, phase)
        allowed = np.zeros((n, n), dtype=bool)
        for s in slots:
            allowed[ri[s], col[s]] = True
        rows_ok = np.zeros(n, dtype=bool)
        for r in rows:
            rows_ok[r - 1] = True
        assert not (A & ~allowed).any(), (label, phase, np.argwhere(A & ~allowed)[:5].tolist())
        assert not (B & ~rows_ok).any(), (label, phase, np.flatnonzero(B & ~rows_ok)[:5].tolist())
        seen += int(A.sum()) + int(B.sum())
    return seen


def test_only_the_declared_stamps_change_inside_a_newton_loop():
    """solve_blocks = 2 solves a matrix block only in the first solve of a
    simulatorSolve loop (core/simulator.c:45-69) if no A slot / right-hand-side
    row the generated code lists as "linearize can change" lies in it.  The
    oracle - pinned to the reference - watches its own stamps between consecutive
    solves of every loop of every golden circuit and of the IBIS links: whatever
    changed must be on that list."""
    import workloads
    total = 0
    nonlinear = 0
    for name, (build, an, _checks) in sorted(circuits.all_cases().items()):
        seen = _check_newton_stamps(build().netlist(), 0, an, name)
        total += 1
        nonlinear += seen > 0
    for cfg, picks in (("cfg4", (0, 5)), ("cfg5", (3,)), ("cfg2", (7,)), ("cfg3", (1,))):
        w = workloads.make(cfg, 16)
        for i in picks:
            nl, vi_set = w.instance_netlist(i)
            seen = _check_newton_stamps(nl, vi_set, ("tran", w.tstep, w.tstop), "%s[%d]" % (cfg, i))
            assert seen > 0
            total += 1
            nonlinear += 1
    assert total >= 30 and nonlinear >= 10


# ---- T-line history search: compile tlineLocateFn for the HOST and compare it with
# ---- the reference's record-by-record walk on random rings ---------------------
def _walk_like_the_reference(times, lo, count, hint, tp):
    """math/history_interp.c:42-64 + core/history.c:36-59 restated: walk from the
    hint record to THE record with t_j <= tp < t_(j+1) (or the newest one with
    t_j <= tp); past the newest record the last two are used (extrapolation).
    Returns (landing record, earlier, later) in record numbers, or None when the
    records needed have left the ring."""
    i = hint if lo <= hint < count else lo
    while i > lo and times[i] > tp:
        i -= 1
    if times[i] > tp:
        return None
    while i + 1 < count and times[i + 1] <= tp:
        i += 1
    if i + 1 < count:
        return i, i, i + 1
    if i - 1 < lo:
        return None
    return i, i - 1, i


def test_tline_history_search_lands_where_the_reference_walk_lands(tmp_path):
    """The galloping + binary search of the kernel (tlineLocateFn) returns the ring
    positions of the same two records as the reference's walk, for any hint, with
    wrapped rings, delayed times before the oldest and past the newest record."""
    import ctypes as C
    import subprocess
    dev = open(os.path.join(ROOT, "eispice_b200", "csrc", "simcu_device.cuh")).read()
    start = dev.index("DEV int tlineWrap(int p, int R)")
    end = dev.index("DEV double tlineInterp(")
    code = """
#define DEV static inline
#define DEVFN static
struct int4 { int x, y, z, w; };
%s
extern "C" void locate(const double *times, int R, int count, int hint, double tp, int *out)
{
	const int4 r = tlineLocateFn(times, R, count, hint, tp);
	out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
""" % dev[start:end]
    cpp, so = tmp_path / "loc.cpp", tmp_path / "loc.so"
    cpp.write_text(code)
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", str(so), str(cpp)], check=True)
    lib = C.CDLL(str(so))
    lib.locate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
    rng = np.random.default_rng(9)
    found = gone = 0
    for trial in range(300):
        R = int(rng.choice([4, 7, 16, 33, 64]))
        count = int(rng.integers(1, 4 * R))
        # strictly increasing accepted times with bursts of tiny steps
        steps = 10 ** rng.uniform(-15, -10, count)
        t = np.concatenate(([0.0], np.cumsum(steps)))[:count]
        ring = np.full(R, np.nan)
        lo = max(0, count - R)
        for q in range(lo, count):
            ring[q % R] = t[q]
        for _ in range(40):
            hint = int(rng.integers(-2, count + 3))
            u = rng.random()
            if u < 0.1:
                tp = float(t[int(rng.integers(lo, count))])            # exactly on a record
            elif u < 0.2:
                tp = float(t[count - 1] + 10 ** rng.uniform(-14, -10))   # past the newest
            else:
                tp = float(rng.uniform(0.0 if lo == 0 else t[lo] * 0.9, t[count - 1] * 1.05))
            want = _walk_like_the_reference(t, lo, count, hint, tp)
            out = (C.c_int * 4)()
            lib.locate(ring.ctypes.data, R, count, hint, tp, out)
            if want is None:
                assert out[3] == 0, (trial, R, count, hint, tp, list(out))
                gone += 1
            else:
                land, a, b = want
                assert out[3] != 0 and out[0] == land, (trial, R, count, hint, tp, list(out), want)
                assert (out[1], out[2]) == (a % R, b % R)
                assert out[3] - 1 == land % R
                found += 1
    assert found > 5000 and gone > 50


def test_break_angle_shortcut_decides_like_the_two_atan2(tmp_path):
    """checkbreak.c:57-66 compares two atan2 angles in raw SI units; the kernel's
    cbAngleBreak settles most calls from certified bounds on the arguments and
    from single-precision angles, and falls back to the reference's expression
    otherwise.  Host build, 23 million argument pairs incl. zeros,
    infinities, NaNs, negative and equal time differences: same decision."""
    import ctypes as C
    import subprocess
    dev = open(os.path.join(ROOT, "eispice_b200", "csrc", "simcu_device.cuh")).read()
    start = dev.index("DEVFN int cbAngleBreak(")
    end = dev.index("/* state (6 live doubles of the 9 slots)")
    code = """
#include <cmath>
#define DEV static inline
#define DEVFN static
#define HUGE_VAL (__builtin_huge_val())
static double simcuAtan2(double y, double x) { return (y == 0.0 && x > 0.0) ? y : atan2(y, x); }
%s
extern "C" long check(long n, const double *dy0, const double *dx0, const double *dy1, const double *dx1,
		double maxAngle, long *breaks)
{
	long bad = 0, nb = 0;
	for (long k = 0; k < n; k++) {
		const int want = (fabs(simcuAtan2(dy0[k], dx0[k]) - simcuAtan2(dy1[k], dx1[k])) > maxAngle) ? 1 : 0;
		const int got = cbAngleBreak(dy0[k], dx0[k], dy1[k], dx1[k], maxAngle);
		bad += (want != got);
		nb += want;
	}
	*breaks = nb;
	return bad;
}
""" % dev[start:end]
    cpp, so = tmp_path / "ang.cpp", tmp_path / "ang.so"
    cpp.write_text(code)
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(cpp)], check=True)
    lib = C.CDLL(str(so))
    lib.check.restype = C.c_long
    lib.check.argtypes = [C.c_long] + [C.c_void_p] * 4 + [C.c_double, C.c_void_p]
    rng = np.random.default_rng(21)
    n = 1 << 19
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, 1e-300, 1e300, 1e-11, 1.0])

    def sample(kind):
        mag = 10.0 ** rng.uniform(-22, 3, n)
        dy = mag * rng.choice([-1.0, 1.0], n)
        dy[rng.random(n) < 0.25] = 0.0
        dx = 10.0 ** rng.uniform(-16, -8, n)
        if kind == 1:                       # near the threshold: |dy| ~ dx
            dy = dx * rng.uniform(-4, 4, n)
        if kind == 2:                       # special values sprinkled in
            hit = rng.random(n) < 0.3
            dy[hit] = rng.choice(special, hit.sum())
            hit = rng.random(n) < 0.3
            dx[hit] = rng.choice(np.concatenate((special, -special[6:])), hit.sum())
        return np.ascontiguousarray(dy), np.ascontiguousarray(dx)

    total_breaks = 0
    for angle in (math.pi / 3, 0.95, 1.09, 0.5, 1.3):
        for k0 in range(3):
            for k1 in range(3):
                dy0, dx0 = sample(k0)
                dy1, dx1 = sample(k1)
                nb = C.c_long()
                bad = lib.check(n, dy0.ctypes.data, dx0.ctypes.data, dy1.ctypes.data, dx1.ctypes.data,
                                angle, C.byref(nb))
                assert bad == 0, (angle, k0, k1, bad)
                total_breaks += nb.value
    assert total_breaks > 1 << 19
