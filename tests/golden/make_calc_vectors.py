#!/usr/bin/env python
"""Golden vectors for the expression engine, produced by the reference's
calculon (calcSolve / calcDiff through oracle/_ref).  Build container only."""
import json
import math
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import backends  # noqa: E402
from test_calc_parity import random_expr  # noqa: E402

NAMES = ["v(a)", "v(b)", "i(Vx)", "time"]
FIXED = [
    "-2^2", "2^3^2", "2(3)^2", "1/0", "0/0", "-1/0", "v(a)^2*i(Vx)", "if(v(a)>1)*3",
    "+2", "1Meg+1m+1M", "2*-3", "v(a)(3)", "-v(a)(3)^2", "if((v(a)>1)&&(i(Vx)<1))",
    "if((v(a)>1)||(i(Vx)<1))", "if(v(a)>1+2)", "if(!v(a)>1)", "2exp(1)", "1e3k", "1mil",
    "V(a)+I(Vx)", "2 3", "sin(1)(2)", "2^-1^2", "3-2-1", "8/2/2", "if(1)", "uramp(0)",
    "u(0)", "u(v(a)-v(a))", "uramp(v(a)-v(a))", "ln(-2)", "log(100)", "sqrt(v(a))",
    "acosh(2)", "acos(0.5)", "asin(0.5)", "asinh(v(b))", "atanh(0.3)", "cosh(v(a))",
    "tan(v(b))", "abs(v(b))", "abs(0)", "v(a)^v(b)", "0^2", "v(a,b)*2", "v(a, b)",
    "time*2", "1.5e-3u", "3.3T", "2G", "1f+1p+1n+1u", "if((v(a))>v(b))", "if(((v(a)>0)))",
    "if((v(a)>0)&&(v(b)>0)&&(time>0))", "if(((v(a)>0)&&(v(b)>0))&&(time>0))",
    "if(v(a)==v(a))", "if(v(a)!=v(a))", "if(v(a)>=v(a))", "if(v(a)<=v(b))",
    "4.739057e-04 * (uramp( v(a,0) --5.060000e+00)) + sqrt(v(b)) / sinh(i(Vx))",
    "-1m*v(a) + 1m*v(a)^3", "sin(2*3.14159*100e6*time)", "foo", "v(a", "1..2", "",
    "(v(a)", "v(a))", "2**3", "1e", "1e+", ".5", "5.", "exp(-v(a)/2.586495e-02)",
]


def main():
    rng = random.Random(1234)
    exprs = list(FIXED) + [random_expr(rng) for _ in range(400)]
    cases = []
    for e in exprs:
        vals = [round(rng.uniform(-2, 2), 6) for _ in NAMES]
        if e in FIXED[:80]:
            vals = [2.0, 0.5, 0.75, 1e-9] if rng.random() < 0.7 else vals
        rc, val, der = backends.eval_expr("ref", e, NAMES, vals)
        case = {"expr": e, "values": vals, "ok": rc == 0}
        if rc == 0:
            case["value"] = val
            case["derivs"] = [None if math.isnan(d) else d for d in der]
        cases.append(case)
    json.dump({"names": NAMES, "cases": cases},
              open(os.path.join(HERE, "calc_vectors.json"), "w"), indent=0)
    print("wrote", len(cases), "cases;", sum(c["ok"] for c in cases), "valid")


if __name__ == "__main__":
    main()
