"""IBIS pin composition, mirroring the reference's ``module/ibis_device.py``
(Receiver :39-82, Driver :84-150, Package :152-177, Pin :179-251).

The IBIS *parser* and K-table generation are host pre-processing and out of
scope; models arrive here as plain dictionaries of
tables (see tests/golden/ibis_tables.npz for the reference's own fixture):

  model = {"model_type": "output", "voltage_range": {speed: V},
           "c_comp": {speed: F}, "gnd_clamp": {speed: table},
           "power_clamp": {...}, "pullup": {...}, "pulldown": {...},
           "pullup_k": {direction: {speed: table}}, "pulldown_k": {...}}
  pin   = {"R": ohm, "L": henry, "C": farad}

``speed`` / ``direction`` may be LISTS: every VI device then carries one table
per (speed, direction) combination -- a table set -- and ``set`` selects among
them per instance in ``Circuit.tran_batch`` (IBIS corner sweeps).
"""
from . import device, subckt, waveform

Typical, Minimum, Maximum = "typ", "min", "max"
Rising, Falling = "rising", "falling"
Output, Input = "output", "input"


def _aslist(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


def corner_sets(speed, direction=None):
    """[(speed, direction), ...] in table-set order."""
    if direction is None:
        return [(s, None) for s in _aslist(speed)]
    return [(s, d) for s in _aslist(speed) for d in _aslist(direction)]


def _vi(tab, sets):
    tabs = [waveform.PWL(tab[s]) for s, _ in sets]
    return tabs if len(tabs) > 1 else tabs[0]


def _ta(tab, sets):
    tabs = [waveform.PWL(tab[d][s]) for s, d in sets]
    return tabs if len(tabs) > 1 else tabs[0]


class Receiver(subckt.Subckt):
    def __init__(self, iNode, pwrNode, gndNode, model, speed=Typical, sets=None):
        sets = sets or corner_sets(speed)
        s0 = sets[0][0]
        if "c_comp" in model:
            self.C_comp = device.C(iNode, gndNode, model["c_comp"][s0])
        if "gnd_clamp" in model:
            self.VI_gnd = device.VI(iNode, gndNode, _vi(model["gnd_clamp"], sets))
        if "power_clamp" in model:
            self.VI_pwr = device.VI(pwrNode, iNode, _vi(model["power_clamp"], sets))


class Driver(subckt.Subckt):
    def __init__(self, oNode, pwrNode, gndNode, model, speed=Typical,
                 direction=Rising, sets=None):
        sets = sets or corner_sets(speed, direction)
        s0 = sets[0][0]
        if "c_comp" in model:
            self.C_comp = device.C(oNode, gndNode, model["c_comp"][s0])
        if "gnd_clamp" in model:
            self.VI_gnd = device.VI(oNode, gndNode, _vi(model["gnd_clamp"], sets))
        if "power_clamp" in model:
            self.VI_pwr = device.VI(pwrNode, oNode, _vi(model["power_clamp"], sets))
        if "pullup" in model:
            self.VI_pu = device.VI(pwrNode, oNode, _vi(model["pullup"], sets),
                                   _ta(model["pullup_k"], sets))
        if "pulldown" in model:
            self.VI_pd = device.VI(oNode, gndNode, _vi(model["pulldown"], sets),
                                   _ta(model["pulldown_k"], sets))


class Package(subckt.Subckt):
    def __init__(self, pinNode, dieNode, gndNode, pin):
        self.Cpkg = device.C(pinNode, gndNode, pin["C"])
        self.Rpkg = device.R(self.node("node"), pinNode, pin["R"])
        self.Lpkg = device.L(dieNode, self.node("node"), pin["L"])


class Pin(subckt.Subckt):
    """Buffer + package of one pin; ``io`` picks Receiver or Driver for I/O
    models exactly like the reference (ibis_device.py:220-234)."""

    def __init__(self, model, pin, node, speed=Typical, direction=Rising,
                 io=Output):
        vcc = self.node("vcc")
        die = self.node("die")
        s0 = _aslist(speed)[0]
        # one table set per (speed, direction) pair on EVERY VI device of the
        # pin, so a single per-instance index selects a consistent corner
        sets = corner_sets(speed, direction)
        self.Vcc = device.V(vcc, 0, model["voltage_range"][s0])
        mtype = model["model_type"]
        if mtype == "input" or (mtype == "i/o" and io == Input):
            self.Buffer = Receiver(die, vcc, 0, model, speed, sets=sets)
        elif mtype in ("output", "3-state", "i/o"):
            self.Buffer = Driver(die, vcc, 0, model, speed, direction, sets=sets)
        else:
            raise RuntimeError("Unsupported Model Type %s" % mtype)
        self.Package = Package(node, die, 0, pin)
