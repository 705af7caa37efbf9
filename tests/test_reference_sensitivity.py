"""What the reference itself does under another pivot order - the facts the GPU
parity gates are built on (CPU only; needs the unmodified reference build
oracle/_ref, which travels to the GPU box prebuilt).

* The oracle's attempt / linearize logs are self-consistent (they drive the
  GPU replay tests).
* SuperLU's perm_c is one fixed permutation per circuit, perm_r takes a handful
  of values (oracle/ref_probe.c reads them out of the reference's own dgssvx).
* The oracle REPLAYING SuperLU's pivot sequence reproduces the reference's own
  outcome - including its abort ("Timestep too small" at 1.1e-13 s) on every
  typ/falling instance of the cfg5 sweep, which the oracle's natural-order
  partial pivoting (and the GPU's own sequence) complete.
* On the IBIS links the reference under two pivot orders (its COLAMD order vs
  the oracle's natural order - the same algorithm) does NOT agree to 1 x its own
  SPICE tolerance: the accepted-step sequence is decided by Newton noise.  The
  adaptive GPU gate (tests/test_gpu_fullsize.py) therefore compares
  distributions; same-sequence parity is tests/test_gpu_replay.py.
"""
import multiprocessing as mp

import numpy as np
import pytest

import backends
import parity
import parity_workers as pw
import workloads

needs_ref = pytest.mark.skipif(not backends.ref_available(), reason="oracle/_ref not built")


@pytest.mark.parametrize("cfg,idx", [("cfg1", 0), ("cfg2", 7), ("cfg4", 3), ("cfg5", 2)])
def test_attempt_and_linearize_logs_are_consistent(cfg, idx):
    w = workloads.make(cfg, 24)
    nl, s = w.instance_netlist(idx)
    o = backends.OracleSim(nl, s)
    o.log_attempts()
    d, _ = o.tran(w.tstep, w.tstop)
    t, f = o.attempt_log()
    lin = o.linearize_log()
    st = o.stats()
    outcome, order, solves = (f >> 4) & 15, f & 15, f >> 8
    assert np.array_equal(t[outcome == 0], d[1:, 0])              # accepted attempts = result rows
    assert (outcome == 1).sum() == st["rejected_lte"] and (outcome == 2).sum() == st["rejected_nc"]
    assert solves.sum() == st["factorizations"] and set(np.unique(order)) <= {1, 2}
    # one linearize pass per solve, except the last solve of a pass that hit the limit
    total = st["factorizations"] + st["op_factorizations"]
    assert total - (outcome == 2).sum() - 1 <= len(lin) <= total
    nonlinear = sum(1 for dev in nl if dev[0] in ("VI", "B"))
    assert lin.max(initial=0) < (1 << max(1, nonlinear))


@needs_ref
@pytest.mark.parametrize("cfg", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_superlu_permutations_and_their_replay(cfg):
    w = workloads.make(cfg, 24)
    idx = 3
    nl, s = w.instance_netlist(idx)
    pc, pr = backends.superlu_sequences(nl, s, w.tstep, w.tstop)
    n = pc.shape[1]
    assert len(pc) >= 2 and all(sorted(x) == list(range(n)) for x in pc[:4])
    assert all(sorted(x) == list(range(n)) for x in pr[:4])
    assert len({tuple(x) for x in pc}) == 1                        # COLAMD sees the pattern only
    assert len({tuple(x) for x in pr}) <= 8
    dr, vr = backends.RefSim(nl, s).tran(w.tstep, w.tstop)
    o = backends.OracleSim(nl, s)
    o.set_pivots(0, pc[0], pr[0])
    o.set_pivots(1, pc[1], pr[1])
    do, vo = o.tran(w.tstep, w.tstop)
    a, b = parity.drop_degenerate(do, w.tstep), parity.drop_degenerate(dr, w.tstep)
    assert parity.grid_error(a[:, 0], a, b[:, 0], b, vo, w.tstep, w.tstop, current_floor=1e-3) < 5.0


@needs_ref
def test_reference_aborts_a_sixth_of_cfg5_under_its_own_pivot_order():
    w = workloads.make("cfg5", 131072)
    backends.oracle_lib().eoQuiet(1)
    aborted = completed = 0
    for idx in (1, 7, 13, 0, 5):                    # corner = idx % 6; 1 = typ / falling
        nl, s = w.instance_netlist(idx)
        try:
            backends.RefSim(nl, s).tran(w.tstep, w.tstop)
            ref_ok = True
        except RuntimeError:
            ref_ok = False
        assert ref_ok == (s != 1), "corner %d" % s
        do, _ = backends.OracleSim(nl, s).tran(w.tstep, w.tstop)      # natural order: completes
        assert do[-1, 0] == w.tstop
        if not ref_ok:
            # the oracle under SuperLU's own sequence aborts exactly like the reference
            pc, pr = _sequences_of_failed_run(nl, s, w)
            o = backends.OracleSim(nl, s)
            o.set_pivots(0, pc[0], pr[0])
            o.set_pivots(1, pc[1], pr[1])
            with pytest.raises(RuntimeError):
                o.tran(w.tstep, w.tstop)
            aborted += 1
        else:
            completed += 1
    backends.oracle_lib().eoQuiet(0)
    assert aborted == 3 and completed == 2


def _sequences_of_failed_run(nl, s, w):
    """superlu_sequences also works for a run the reference aborts."""
    return backends.superlu_sequences(nl, s, w.tstep, w.tstop, max_records=8)


@needs_ref
@pytest.mark.parametrize("cfg,count", [("cfg4", 48), ("cfg5", 24)])
def test_reference_under_two_pivot_orders_misses_its_own_tolerance(cfg, count):
    """reference (COLAMD order) vs oracle (natural order), same algorithm, same
    instances: a visible share of the sweep differs by more than 1 x
    reltol*max|x| + vntol/abstol - with or without the current floor of
    tests/parity.py.  This is why the adaptive GPU gate is distributional."""
    w = workloads.make(cfg)
    pick = np.sort(np.random.default_rng(4).choice(w.M, count, replace=False))
    with mp.get_context("spawn").Pool(8) as pool:
        nat = pool.map(pw.gridded, [(cfg, w.M, int(i), "oracle") for i in pick])
        ref = pool.map(pw.gridded, [(cfg, w.M, int(i), "ref") for i in pick])
    nl, _ = w.instance_netlist(0)
    names = backends.OracleSim(nl).op()[1][1:]
    both = [j for j in range(count) if nat[j] is not None and ref[j] is not None]
    assert len(both) >= count // 2
    e = np.array([pw.spice_ratio(ref[j], nat[j], names) for j in both])
    e_floor = np.array([pw.spice_ratio(ref[j], nat[j], names, 1e-3) for j in both])
    print("%s: reference vs oracle, %d instances: %.2f within 1 x (max %.3g); with the current floor "
          "%.2f within 1 x (max %.3g)" % (cfg, len(both), (e <= 1).mean(), e.max(),
                                          (e_floor <= 1).mean(), e_floor.max()))
    assert (e > 1.0).mean() >= 0.1 and e.max() > 2.0
    assert (e_floor > 1.0).mean() >= 0.04
