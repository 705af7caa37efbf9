"""IBIS switching-coefficient (K) table generation.

eispice_b200.ibis_driver builds the reference's K-extraction circuits
(module/ibis_driver.py) from the reference's own fixture (V-t tables parsed by
the unmodified reference parser, tests/golden/ibis_vt.npz) and runs them
 - on the CPU oracle: every K table of every driver model, direction and speed
   must reproduce the table the reference computed at parse time
   (tests/golden/ibis_tables.npz) point for point;
 - on the GPU (marked gpu): same tables through the CUDA engine.
"""
import numpy as np
import pytest

import backends
import circuits
import parity
from eispice_b200 import ibis_driver

DRIVERS = ["test_out_1", "test_out_2", "test_out_3"]      # double-wave, single->ramp, ramp
COMBOS = [(d, s) for d in ("rising", "falling") for s in ("typ", "min", "max")]


def _oracle_runner(cct, tstep, tstop):
    return backends.OracleSim(cct.netlist()).tran(tstep, tstop)


@pytest.mark.parametrize("name", DRIVERS)
def test_k_tables_on_the_oracle_match_the_reference(name):
    models, _ = circuits.ibis_fixture()
    m = models[name]
    method = ibis_driver.k_circuit(m, "rising", "typ")[0].title
    assert method == ("Double Waveform" if name == "test_out_1" else "Ramp")
    for direction, speed in COMBOS:
        ku, kd = ibis_driver.calc_driver_k(m, direction, speed, runner=_oracle_runner)
        for mine, gold in ((ku, m["pullup_k"][direction][speed]),
                           (kd, m["pulldown_k"][direction][speed])):
            assert mine.shape == gold.shape, (name, direction, speed)
            frac, worst = parity.tight_fraction(mine, gold)
            assert frac == 1.0, (name, direction, speed, worst)


def test_k_circuit_mirrors_the_reference_topology():
    models, _ = circuits.ibis_fixture()
    cct, tstep, tstop = ibis_driver.double_wave_circuit(models["test_out_1"], "rising", "typ")
    refs = [d[1] for d in cct.netlist()]
    assert refs == ["Vcc", "Xwv0", "Xpu0", "Xpd0", "Vmeas0", "Cc0", "Rfix0", "Vfix0", "Xcu0", "Xcd0",
                    "Xwv1", "Xpu1", "Xpd1", "Vmeas1", "Cc1", "Rfix1", "Vfix1", "Xcu1", "Xcd1",
                    "Bkdn", "Bkun"]
    w = models["test_out_1"]["rising_waveform"]
    assert tstop == max(w[0]["data"]["typ"][-1][0], w[1]["data"]["typ"][-1][0])
    assert np.isclose(tstep * 10, min(w[0]["data"]["typ"][1][0] - w[0]["data"]["typ"][0][0],
                                      w[1]["data"]["typ"][1][0] - w[1]["data"]["typ"][0][0]))


@pytest.mark.gpu
@pytest.mark.parametrize("name,direction,speed", [("test_out_1", "rising", "typ"),
                                                  ("test_out_1", "falling", "max"),
                                                  ("test_out_3", "rising", "min"),
                                                  ("test_out_3", "falling", "typ")])
def test_k_tables_on_the_gpu(name, direction, speed):
    models, _ = circuits.ibis_fixture()
    m = models[name]
    ku, kd = ibis_driver.calc_driver_k(m, direction, speed)            # CUDA engine
    for mine, gold in ((ku, m["pullup_k"][direction][speed]), (kd, m["pulldown_k"][direction][speed])):
        # K(t) is what the buffer uses: compare as functions of time at the
        # reference's reltol, and point for point where the step sequences agree
        grid = np.linspace(gold[0, 0], gold[-1, 0], 2000)
        a = np.interp(grid, mine[:, 0], mine[:, 1])
        b = np.interp(grid, gold[:, 0], gold[:, 1])
        assert np.abs(a - b).max() <= parity.RELTOL * max(1.0, np.abs(b).max())
        if mine.shape == gold.shape:
            frac, worst = parity.tight_fraction(mine, gold)
            assert frac >= 0.99, (frac, worst)


@pytest.mark.gpu
def test_link_with_gpu_generated_k_tables():
    """End to end: K tables generated on the GPU drive the cfg4 link; the far-end
    waveform equals the one obtained with the reference's K tables."""
    import copy
    models, pins = circuits.ibis_fixture()
    mine = copy.copy(models["test_out_1"])
    ku, kd = ibis_driver.calc_driver_k(mine, "rising", "typ")
    mine["pullup_k"] = {"rising": {"typ": ku}}
    mine["pulldown_k"] = {"rising": {"typ": kd}}
    import eispice_b200 as eispice
    from eispice_b200 import ibis_device

    def link(model):
        cct = circuits._new("link")
        cct.Driver = ibis_device.Pin(model, pins['2'], 'vsx', "typ", "rising")
        cct.Vmeas = eispice.V('vsx', 'vs', 0)
        cct.Rt = eispice.R('vs', 'vi', '33.2')
        cct.Tg = eispice.T('vi', 0, 'vo', 0, 50, '2n')
        cct.Receiver = ibis_device.Pin(models[pins['1']["model"]], pins['1'], 'vo', "typ", "rising")
        cct.tran('0.01n', '15n')
        return cct
    a, b = link(mine), link(models["test_out_1"])
    grid = np.linspace(0, 15e-9, 1501)
    va = np.interp(grid, a.t, a.results[:, a.variables.index("v(vo)")])
    vb = np.interp(grid, b.t, b.results[:, b.variables.index("v(vo)")])
    assert np.abs(va - vb).max() < parity.RELTOL * np.abs(vb).max() + parity.VNTOL
    assert a.check_v('vo', 2.942970279e-01, '4.7n') or abs(a.v['vo']('4.7n') - 2.942970279e-01) < 1e-3


def test_merged_k_circuits_expand_to_the_single_circuits():
    """Batched K-table generation (CPU side): the extraction circuits of all
    (model, direction, speed) combinations group into a few topologies; merging a
    group into one netlist with table sets + per-instance arrays and expanding
    instance j again gives exactly job j's own netlist."""
    models, _ = circuits.ibis_fixture()
    jobs = []
    for name in DRIVERS:
        for direction, speed in COMBOS:
            cct, tstep, tstop = ibis_driver.k_circuit(models[name], direction, speed)
            jobs.append((name, direction, speed, cct.netlist(), tstep, tstop))
    groups = {}
    for job in jobs:
        groups.setdefault(ibis_driver._signature(job[3]), []).append(job)
    assert 1 <= len(groups) <= 4 and sum(len(g) for g in groups.values()) == 18
    for group in groups.values():
        merged, params = ibis_driver.merge_netlists([g[3] for g in group])
        for j, job in enumerate(group):
            mine = ibis_driver.instance_netlist(merged, params, j)
            assert len(mine) == len(job[3])
            for a, b in zip(mine, job[3]):
                assert a[:4] == b[:4]
                if a[0] in ("R", "C", "L"):
                    assert float(a[4]) == float(b[4])
                elif a[0] in ("V", "I"):
                    assert float(a[5]) == float(b[5])
                    if b[6] is not None and b[6][0] in ("l", "c"):
                        assert np.array_equal(np.asarray(a[6][1], float), np.asarray(b[6][1], float))
                    elif b[6] is not None:
                        x, y = np.array(a[6][1], float), np.array(b[6][1], float)
                        assert ((x == y) | (np.isinf(x) & np.isinf(y))).all()
                elif a[0] == "VI":
                    assert np.array_equal(np.asarray(a[4][0][1], float), np.asarray(b[4][0][1], float))
                    assert (a[5] is None) == (b[5] is None)
        # instance j of the merged netlist simulates like job j (oracle, one instance)
        j = len(group) - 1
        do, vo = backends.OracleSim(ibis_driver.instance_netlist(merged, params, j)).tran(group[j][4], group[j][5])
        dj, vj = backends.OracleSim(group[j][3]).tran(group[j][4], group[j][5])
        assert vo == vj and np.array_equal(do, dj)


@pytest.mark.gpu
def test_k_tables_batched_on_the_gpu():
    """All 18 K-table transients of the fixture's three driver models as instances
    of one batch per extraction topology (per-instance analysis windows, waveform
    and V-I table sets): every table equals the reference's as a function of
    time at the reference's reltol."""
    import time
    models, _ = circuits.ibis_fixture()
    drivers = {name: models[name] for name in DRIVERS}
    t0 = time.time()
    tables, info = ibis_driver.driver_k_tables_batched(drivers)
    cold = time.time() - t0
    t0 = time.time()
    tables, info = ibis_driver.driver_k_tables_batched(drivers)
    warm = time.time() - t0
    assert info["instances"] == 18 and info["failed"] == 0 and info["batches"] <= 4
    for name in DRIVERS:
        for direction, speed in COMBOS:
            for kind in ("pullup_k", "pulldown_k"):
                mine, gold = tables[name][kind][direction][speed], models[name][kind][direction][speed]
                grid = np.linspace(gold[0, 0], gold[-1, 0], 2000)
                a = np.interp(grid, mine[:, 0], mine[:, 1])
                b = np.interp(grid, gold[:, 0], gold[:, 1])
                assert np.abs(a - b).max() <= parity.RELTOL * max(1.0, np.abs(b).max()), (name, direction, speed, kind)
    print("K tables of %d transients in %d batches: %.2f s cold (NVRTC), %.3f s warm; reference: 0.22 s on one core"
          % (info["instances"], info["batches"], cold, warm))
