"""Expression-engine parity: oracle restatement of calculon vs golden vectors
produced by the reference's calcSolve/calcDiff (tests/golden/calc_vectors.json,
made by tests/golden/make_calc_vectors.py), and live differential fuzzing when
oracle/_ref is available.  Semantics pinned: SURVEY.md 8a row 12."""
import json
import math
import os
import random

import pytest

import backends

VEC = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "calc_vectors.json")))


def _same(a, b):
    if a is None or b is None:
        return a is None and b is None
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    if math.isinf(a) or math.isinf(b):
        return a == b
    return abs(a - b) <= 1e-12 * max(abs(a), abs(b)) + 1e-300


def test_golden_vectors():
    names = VEC["names"]
    for case in VEC["cases"]:
        rc, val, der = backends.eval_expr("oracle", case["expr"], names, case["values"])
        assert (rc == 0) == case["ok"], case["expr"]
        if rc == 0:
            assert _same(val, case["value"]), (case["expr"], val, case["value"])
            for d, g in zip(der, case["derivs"]):
                assert _same(None if math.isnan(d) else d, g), (case["expr"], der, case["derivs"])


def random_expr(rng, depth=0):
    atoms = ["v(a)", "v(b)", "i(Vx)", "time", "2", "0.5", "1m", "3k", "1e-3", "v(a,b)"]
    funcs = ["abs", "sin", "cos", "exp", "sqrt", "uramp", "u", "atan", "sinh", "ln", "log", "tan"]
    if depth > 3 or rng.random() < 0.25:
        return rng.choice(atoms)
    r = rng.random()
    if r < 0.45:
        return "%s%s%s" % (random_expr(rng, depth + 1), rng.choice(["+", "-", "*", "/", "^"]),
                           random_expr(rng, depth + 1))
    if r < 0.6:
        return "(%s)" % random_expr(rng, depth + 1)
    if r < 0.7:
        return "-%s" % random_expr(rng, depth + 1)
    if r < 0.8:
        return "%s(%s)" % (random_expr(rng, depth + 1), random_expr(rng, depth + 1))
    if r < 0.9:
        return "if(%s%s%s)" % (random_expr(rng, depth + 1), rng.choice([">", "<", ">=", "<=", "==", "!="]),
                               random_expr(rng, depth + 1))
    return "%s(%s)" % (rng.choice(funcs), random_expr(rng, depth + 1))


@pytest.mark.skipif(not backends.ref_available(), reason="oracle/_ref not built")
def test_fuzz_against_reference_calculon():
    rng = random.Random(20260)
    names = ["v(a)", "v(b)", "i(Vx)", "time"]
    checked = 0
    for _ in range(1500):
        expr = random_expr(rng)
        vals = [rng.uniform(-2, 2) for _ in names]
        r_rc, r_val, r_der = backends.eval_expr("ref", expr, names, vals)
        o_rc, o_val, o_der = backends.eval_expr("oracle", expr, names, vals)
        assert (r_rc == 0) == (o_rc == 0), expr
        if r_rc == 0:
            assert _same(o_val, r_val), (expr, vals, o_val, r_val)
            for a, b in zip(o_der, r_der):
                assert _same(a, b), (expr, vals, o_der, r_der)
            checked += 1
    assert checked > 500
