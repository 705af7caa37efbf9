"""eispice_b200 -- B200-native batched transient engine behind eispice's API.

Only the hot path of eispice is here: netlist flattening,
device stamps, Newton loop, numeric LU and step control, run on the GPU over
many instances of one topology.  The classes below keep the reference's names
and call shapes (module/__init__.py:52-61) so eispice scripts port unchanged.
"""
from . import units, subckt                                         # noqa: F401
from .subckt import Subckt                                          # noqa: F401
from .waveform import PWL, PWC, SFFM, Exp, Pulse, Gauss, Sin        # noqa: F401
from .device import (R, C, L, V, I, B, VI, G, E, F, H, T, D, RealC,  # noqa: F401
                     Current, Voltage, Capacitor, Time, GND)
from .circuit import Circuit                                        # noqa: F401
from . import ibis_device                                           # noqa: F401
