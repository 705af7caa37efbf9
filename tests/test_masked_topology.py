"""Per-instance topology parameters: a MAXIMUM topology
whose T-line sections can be bypassed per instance ("Tk.bypass" = 1: the two
ports are joined by an ideal 0 V branch), so that a sweep over the NUMBER of
line sections runs as one batch with one kernel.  The reference has no such
feature; the check is electrical: an instance with n active sections must equal
the reference's / oracle's run of the link BUILT with n sections."""
import numpy as np
import pytest

import backends
import parity
import parity_workers as pw
import workloads
from eispice_b200 import engine


def _masked_oracle(w, i):
    nl, s = w.instance_netlist(i)
    o = backends.OracleSim(nl, s)
    for k in range(1, 9):
        o.set_param("T%d" % k, "bypass", float(w.params["T%d.bypass" % k][i]))
    return o


@pytest.mark.parametrize("i", [0, 2, 5, 7])
def test_bypassed_sections_equal_the_reduced_topology_on_the_oracle(i):
    w = workloads.cfg5_masked(16)
    d1, v1 = _masked_oracle(w, i).tran(w.tstep, w.tstop)
    nl2, s2 = workloads.reduced_link_netlist(w, i)
    runner = backends.RefSim if backends.ref_available() and s2 != 1 else backends.OracleSim
    d2, v2 = runner(nl2, s2).tran(w.tstep, w.tstop)
    g = pw.grid_of(w.tstep, w.tstop)
    for name in w.probes:
        a = np.interp(g, d1[:, 0], d1[:, v1.index(name)])
        b = np.interp(g, d2[:, 0], d2[:, v2.index(name)])
        tol = parity.RELTOL * np.abs(b).max() + (parity.VNTOL if name[0] == "v" else parity.ABSTOL)
        assert np.abs(a - b).max() <= 2.0 * tol, (i, name, float(np.abs(a - b).max() / tol))
    # the bypassed sections carry the through current and no second-port current
    n = int(w.n_active[i])
    if n < 8:
        k = n + 1
        assert np.abs(d1[:, v1.index("i(T%d#in2)" % k)]).max() == 0.0
        # ideal connection: both ports of a bypassed section at the same voltage
        prev = "v(vi)" if k == 1 else "v(n%d)" % (k - 1)
        nxt = "v(vo)" if k == 8 else "v(n%d)" % k
        assert np.abs(d1[:, v1.index(prev)] - d1[:, v1.index(nxt)]).max() < 1e-9


@pytest.mark.gpu
def test_masked_sections_on_the_gpu():
    """One batch, one kernel, every instance with its own number of active
    sections: equal to the oracle with the same flags (same algorithm) and to the
    reduced topology (the reference where it completes, else the oracle)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    engine.set_device(0)
    w = workloads.cfg5_masked(48)
    eng = engine.Engine(w.cct.netlist(), w.M)
    res = eng.tran_batch(w.tstep, w.tstop, 0.0, w.params, w.probes)
    assert (res.status == 0).all(), res.status
    assert res.stats["jitCompiles"] == 1
    steps = eng.counters("accepted")
    # fewer active sections -> fewer delayed reflections to resolve
    assert steps[w.n_active == 1].mean() < steps[w.n_active == 8].mean()
    worst_o = worst_r = 0.0
    for i in (0, 3, 9, 14, 20, 31, 47):
        od, ov = _masked_oracle(w, i).tran(w.tstep, w.tstop)
        nl2, s2 = workloads.reduced_link_netlist(w, i)
        runner = backends.RefSim if backends.ref_available() and s2 != 1 else backends.OracleSim
        rd, rv = runner(nl2, s2).tran(w.tstep, w.tstop)
        for p, name in enumerate(w.probes):
            for data, names, which in ((od, ov, "o"), (rd, rv, "r")):
                d = parity.drop_degenerate(data, w.tstep)
                ref = np.interp(res.grid, d[:, 0], d[:, names.index(name)])
                tol = parity.RELTOL * np.abs(ref).max() + (parity.VNTOL if name[0] == "v" else parity.ABSTOL)
                e = float(np.abs(res.data[i, :, p] - ref).max() / tol)
                if which == "o":
                    worst_o = max(worst_o, e)
                else:
                    worst_r = max(worst_r, e)
    print("masked topology: worst %.3g x tol vs the oracle with the same flags, %.3g x tol vs the reduced topology"
          % (worst_o, worst_r))
    assert worst_o < 5.0 and worst_r < 5.0
    eng.close()
