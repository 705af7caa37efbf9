"""Process-pool workers for the parity tests: CPU oracle / unmodified reference
runs of single sweep instances (test infrastructure only)."""
import numpy as np

_CACHE = {}


def _workload(cfg, M):
    key = (cfg, M)
    if key not in _CACHE:
        import workloads
        _CACHE.clear()
        _CACHE[key] = workloads.make(cfg, M)
    return _CACHE[key]


def grid_of(tstep, tstop):
    g = np.arange(0.0, tstop + 0.5 * tstep, tstep)
    return g[g <= tstop]


def to_grid(data, tstep, tstop):
    """[ngrid, nvar-1]: every unknown linearly interpolated onto the tstep grid."""
    import parity
    d = parity.drop_degenerate(data, tstep)
    g = grid_of(tstep, tstop)
    return np.stack([np.interp(g, d[:, 0], d[:, j]) for j in range(1, d.shape[1])], axis=1)


def oracle_logged(args):
    """Oracle with a fixed pivot sequence (the GPU's) and the attempt log:
    (raw record, log times, log flags, linearize masks), or None if it aborts."""
    cfg, M, idx, piv = args
    import backends
    w = _workload(cfg, M)
    nl, s = w.instance_netlist(idx)
    o = backends.OracleSim(nl, s)
    if piv is not None:
        o.set_pivots(0, piv[0], piv[1])
        o.set_pivots(1, piv[2], piv[3])
    o.log_attempts()
    try:
        d, _ = o.tran(w.tstep, w.tstop)
    except RuntimeError:
        return None
    t, f = o.attempt_log()
    return d, t, f, o.linearize_log()


def gridded(args):
    """(kind, gridded waveforms) of one instance on 'oracle' (partial pivoting,
    natural column order) or 'ref' (the unmodified reference, COLAMD order);
    None if that engine aborts on it."""
    cfg, M, idx, kind = args
    import backends
    w = _workload(cfg, M)
    nl, s = w.instance_netlist(idx)
    try:
        if kind == "ref":
            d, _ = backends.RefSim(nl, s).tran(w.tstep, w.tstop)
        else:
            d, _ = backends.OracleSim(nl, s).tran(w.tstep, w.tstop)
    except RuntimeError:
        return None
    return to_grid(d, w.tstep, w.tstop)


def spice_ratio(a, ref, names, current_floor=0.0):
    """Worst |a - ref| over the grid in units of reltol*max|ref| + vntol/abstol
    per unknown (names without 'time'); see parity.grid_error for the floor."""
    import parity
    scale = np.abs(ref).max(axis=0)
    isI = np.array([n[0] == "i" for n in names])
    if current_floor > 0.0 and isI.any():
        scale = np.where(isI, np.maximum(scale, current_floor * scale[isI].max()), scale)
    tol = parity.RELTOL * scale + np.where(isI, parity.ABSTOL, parity.VNTOL)
    return float((np.abs(a - ref) / tol).max())


def probes_on_grid(args):
    """Probe waveforms of one sweep instance on a given time grid:
    (cfg, M, index, kind 'oracle' | 'ref', probes, grid) -> [ngrid, nprobes] or
    None if that engine aborts on the instance."""
    cfg, M, idx, kind, probes, grid = args
    import backends
    import parity
    w = _workload(cfg, M)
    nl, s = w.instance_netlist(idx)
    try:
        if kind == "ref":
            d, names = backends.RefSim(nl, s).tran(w.tstep, w.tstop)
        else:
            d, names = backends.OracleSim(nl, s).tran(w.tstep, w.tstop)
    except RuntimeError:
        return None
    d = parity.drop_degenerate(d, w.tstep)
    return np.stack([np.interp(grid, d[:, 0], d[:, names.index(p)]) for p in probes], axis=1)
