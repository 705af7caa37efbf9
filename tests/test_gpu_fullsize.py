"""The five BASELINE configurations at their FULL sizes (cfg1 65 536, cfg2
10 000, cfg3 100 000, cfg4 262 144, cfg5 131 072 = one GPU's share of the
1 M-instance north-star run), adaptive stepping, through the C-ABI: 256 random
instances of every sweep are compared with the CPU oracle (the reference's
algorithm with partial pivoting) and with the UNMODIFIED reference, all linearly
interpolated onto the tstep grid, in units of the reference's own SPICE
tolerance reltol*max|x| + vntol/abstol (1e-3, 1e-6 V, 1e-12 A).

Gate: 1 x that tolerance, no current floor, for cfg1 / cfg2 / cfg3.  On the
IBIS links (cfg4, cfg5) the reference ITSELF does not meet 1 x against its own
algorithm under another pivot order (COLAMD column order vs natural order: the
accepted-step sequence is decided by reltol-level Newton noise, SURVEY.md 7
"parity is control-flow parity") - the test measures that spread on the same
instances and requires the engine to be at least as close to the oracle as the
reference is.  Same-sequence parity at 1e-6 / 1e-9 is tests/test_gpu_replay.py.

Flagged instances (status != 0; the batch continues where the reference would
abort the whole run) get their second chance inside tran_batch(rescue=True): a
small batch of their own with other pivots and the deepest history ring.  Every
instance still flagged after that must abort in the reference as well, and
every rescued one is inside the gate."""
import multiprocessing as mp
import os

import numpy as np
import pytest

import backends
import parity_workers as pw
import workloads
from eispice_b200 import engine

pytestmark = pytest.mark.gpu
NSAMPLE = 256


@pytest.fixture(scope="module")
def pool():
    p = mp.get_context("spawn").Pool(min(32, len(os.sched_getaffinity(0))))
    yield p
    p.close()


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("GPU tests need a CUDA device (there is no CPU fallback)")
    engine.set_device(0)


def _ratios(data, refs, probes):
    out = []
    for d, r in zip(data, refs):
        if r is not None:
            out.append(pw.spice_ratio(d, r, probes))
    return np.array(out)


@pytest.mark.parametrize("cfg", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_full_size_sweep_against_oracle_and_reference(cfg, pool):
    w = workloads.make(cfg)
    assert w.M == workloads.DEFAULT_INSTANCES[cfg]
    eng = engine.Engine(w.cct.netlist(), w.M)
    res = eng.tran_batch(w.tstep, w.tstop, 0.0, w.params, w.probes)
    st = res.stats
    ok = res.status == 0
    flagged = np.nonzero(~ok)[0]
    assert ok.mean() > 0.998, "failed instances: %d of %d" % (len(flagged), w.M)
    assert np.isfinite(res.data[ok]).all()
    assert np.isnan(res.data[~ok][:, -1, :]).all()          # flagged: NaN from the failure time on
    grid = res.grid
    pick = np.sort(np.random.default_rng(123).choice(np.nonzero(ok)[0], NSAMPLE, replace=False))
    nat = pool.map(pw.probes_on_grid, [(cfg, w.M, int(i), "oracle", w.probes, grid) for i in pick])
    have_ref = backends.ref_available()
    ref = pool.map(pw.probes_on_grid, [(cfg, w.M, int(i), "ref", w.probes, grid) for i in pick]) \
        if have_ref else [None] * len(pick)
    both = [j for j in range(len(pick)) if nat[j] is not None]
    assert len(both) >= 0.95 * len(pick)
    e_gpu = _ratios([res.data[pick[j]] for j in both], [nat[j] for j in both], w.probes)
    refd = [j for j in both if ref[j] is not None]
    e_ref = _ratios([ref[j] for j in refd], [nat[j] for j in refd], w.probes)
    e_gpu_ref = _ratios([res.data[pick[j]] for j in refd], [ref[j] for j in refd], w.probes)
    print("%s @ %d: %d instances | gpu vs oracle: %.3f within 1x, p90 %.3g, max %.3g | reference vs oracle "
          "(%d run): %.3f within 1x, p90 %.3g, max %.3g | gpu vs reference: %.3f within 1x, max %.3g | flagged %d"
          % (cfg, w.M, len(both), (e_gpu <= 1).mean(), np.quantile(e_gpu, 0.9), e_gpu.max(), len(refd),
             (e_ref <= 1).mean() if len(e_ref) else -1, np.quantile(e_ref, 0.9) if len(e_ref) else -1,
             e_ref.max() if len(e_ref) else -1, (e_gpu_ref <= 1).mean() if len(e_ref) else -1,
             e_gpu_ref.max() if len(e_ref) else -1, len(flagged)))
    if cfg in ("cfg1", "cfg2", "cfg3"):
        assert (e_gpu <= 1.0).all(), "outside 1 x SPICE tolerance: max %.3g" % e_gpu.max()
        if len(e_ref):
            assert (e_gpu_ref <= 1.0).all()
    else:
        # the reference under two pivot orders (its own COLAMD order vs the oracle's
        # natural order) does not meet 1 x on these circuits ...
        if len(e_ref) >= 100:
            assert (e_ref > 1.0).mean() > 0.03, "the reference met 1 x: tighten this gate"
            # ... and the engine is at least as close to the oracle as the reference is
            assert (e_gpu <= 1.0).mean() >= (e_ref <= 1.0).mean() - 0.02
            assert np.quantile(e_gpu, 0.9) <= max(1.0, 1.25 * np.quantile(e_ref, 0.9))
            assert e_gpu.max() <= max(5.0, 2.0 * e_ref.max())
        else:
            assert (e_gpu <= 1.0).mean() >= 0.75 and e_gpu.max() < 50.0

    # ---- flagged instances: second chance, then they must abort in the reference too
    first_pass = flagged
    if len(first_pass):
        res2 = eng.tran_batch(w.tstep, w.tstop, 0.0, w.params, w.probes, rescue=True)
        assert np.array_equal(res2.status == 0, ok) or res2.stats["rescued"] > 0
        still = np.nonzero(res2.status != 0)[0]
        rescued = np.array([i for i in first_pass if res2.status[i] == 0], dtype=int)
        check = np.concatenate([still[:64], rescued[:64]]).astype(int)
        outcome = pool.map(pw.probes_on_grid,
                           [(cfg, w.M, int(i), "ref" if have_ref else "oracle", w.probes, grid) for i in check])
        by = dict(zip(check.tolist(), outcome))
        alive = [int(i) for i in still[:64] if by[int(i)] is not None]
        assert not alive, "still flagged after the rescue, but fine in the reference: %s" % alive
        worst = 0.0
        for i in rescued[:64]:
            if by[int(i)] is not None:
                worst = max(worst, pw.spice_ratio(res2.data[i], by[int(i)], w.probes))
        assert worst <= max(5.0, 2.0 * (e_ref.max() if len(e_ref) else 25.0)), worst
        print("%s: %d flagged in the first pass, %d rescued (worst %.3g x tol vs the reference), %d still "
              "flagged - all of those abort in the reference as well"
              % (cfg, len(first_pass), len(rescued), worst, len(still)))
    eng.close()
