"""Shared parity metrics (used by CPU and GPU tests)."""
import numpy as np

# fixed-step / same-sequence gate
RTOL_TIGHT, ATOL_TIGHT = 1e-6, 1e-9
# adaptive gate = the reference's own SPICE tolerances (control.c:55-57)
RELTOL, VNTOL, ABSTOL = 1e-3, 1e-6, 1e-12


def tight_fraction(a, ref):
    """Fraction of samples with |a-ref| <= 1e-6|ref| + 1e-9, and the worst ratio."""
    err = np.abs(a - ref)
    tol = RTOL_TIGHT * np.abs(ref) + ATOL_TIGHT
    ratio = err / tol
    return float((ratio <= 1.0).mean()), float(ratio.max())


def grid_error(ta, ya, tr, yr, names, tstep, tstop):
    """Worst error after linear interpolation of both waveforms onto the tstep
    grid, in units of reltol*max|x| + vntol/abstol (<= 1 passes)."""
    grid = np.arange(0.0, tstop + 0.5 * tstep, tstep)
    grid = grid[grid <= tstop]
    worst = 0.0
    for j, name in enumerate(names):
        if name == "time":
            continue
        a = np.interp(grid, ta, ya[:, j])
        r = np.interp(grid, tr, yr[:, j])
        tol = RELTOL * np.abs(r).max() + (VNTOL if name[0] == "v" else ABSTOL)
        worst = max(worst, float((np.abs(a - r) / tol).max()))
    return worst


def reference_check(value, sim_value):
    """The reference's own +/-0.01 % doctest criterion (module/circuit.py:241-260)."""
    sign = 1 if value >= 0 else -1
    return value * (1.0 - sign * 1e-4) <= sim_value <= value * (1.0 + sign * 1e-4)
