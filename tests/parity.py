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


def grid_error(ta, ya, tr, yr, names, tstep, tstop, current_floor=0.0):
    """Worst error after linear interpolation of both waveforms onto the tstep
    grid, in units of reltol*max|x| + vntol/abstol (<= 1 passes).

    current_floor > 0 scales the tolerance of a branch current by at least
    current_floor * (largest branch current of the circuit).  The reference's
    Newton test accepts a table device when its current is within reltol of
    itself (math/checklinear.c:39-91) and its LTE control acts on node voltages
    at reltol; the leakage current of a clamp that is off (1e-7 A beside a
    0.15 A driver) is therefore only defined to ~G*reltol*|v| and moves by more
    than reltol of ITSELF between two legitimate pivot orders of the reference
    algorithm (measured: oracle with partial pivoting vs oracle with a fixed
    sequence, cfg4 corner max/falling: 12 x).  The floor keeps such traces from
    deciding the gate when two different accepted-step sequences are compared."""
    grid = np.arange(0.0, tstop + 0.5 * tstep, tstep)
    grid = grid[grid <= tstop]
    worst = 0.0
    imax = 0.0
    if current_floor > 0.0:
        for j, name in enumerate(names):
            if name[0] == "i":
                imax = max(imax, float(np.abs(yr[:, j]).max()))
    for j, name in enumerate(names):
        if name == "time":
            continue
        a = np.interp(grid, ta, ya[:, j])
        r = np.interp(grid, tr, yr[:, j])
        scale = np.abs(r).max()
        if name[0] == "i":
            scale = max(scale, current_floor * imax)
        tol = RELTOL * scale + (VNTOL if name[0] == "v" else ABSTOL)
        worst = max(worst, float((np.abs(a - r) / tol).max()))
    return worst


def reference_check(value, sim_value):
    """The reference's own +/-0.01 % doctest criterion (module/circuit.py:241-260)."""
    sign = 1 if value >= 0 else -1
    return value * (1.0 - sign * 1e-4) <= sim_value <= value * (1.0 + sign * 1e-4)


def drop_degenerate(data, tstep):
    """Remove accepted points that end a step shorter than 1e-9*tstep.

    The reference clips the last step to tstop (core/simulator.c:157-161); when
    the previous point already sits one ulp before tstop that final step is
    ~1e-23 s long, the companion conductance 2C/h is ~1e15 and the branch
    currents of that one sample are rounding noise in the reference itself
    (its own LTE path aborts below tstep*1e-9, :193-195).  Such samples carry
    no information and are excluded from every gate."""
    t = data[:, 0]
    keep = np.ones(len(t), bool)
    keep[1:] = np.diff(t) >= 1e-9 * tstep
    return data[keep]
