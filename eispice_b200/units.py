"""Unit-suffix parsing, mirroring the reference's ``module/units.py``.

``float('10nF') == 1e-08``; fractions such as ``'12nF/2nH'`` are supported.
The arithmetic is kept identical to the reference (``builtin float(value) *
multiplier``, reference module/units.py:66-89) so that netlists built from
strings carry bit-identical parameters.  Multipliers (reference
module/units.py:54-62): T G M k m u n p f -- note ``M`` is mega here, unlike
inside B-source equations where calculon reads ``m``/``M`` as milli.
"""
import builtins
import re

import numpy as np

_NUM = r"[-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?"
_pattern = re.compile(
    r"\s*(?P<vn>" + _NUM + r")(?P<un>\w*)\s*"
    r"(?:/\s*(?P<vd>" + _NUM + r")(?P<ud>\w*))?")

_MULT = (("T", 1e12), ("G", 1e9), ("M", 1e6), ("k", 1e3), ("m", 1e-3),
         ("u", 1e-6), ("n", 1e-9), ("p", 1e-12), ("f", 1e-15))


def _scaled(value, suffix):
    v = builtins.float(value)
    if suffix:
        for letter, mult in _MULT:
            if suffix.startswith(letter):
                return v * mult
    return v


def float(value):  # noqa: A001 - same public name as the reference
    """Float from a number or a string with a unit suffix (None if no match)."""
    if not isinstance(value, str):
        return builtins.float(value)
    m = _pattern.match(value)
    if m is None:
        return None
    num = _scaled(m.group("vn"), m.group("un"))
    if m.group("vd") is not None:
        return num / _scaled(m.group("vd"), m.group("ud"))
    return num


def floatList1D(data):
    """1-D list of numbers / unit strings -> float ndarray."""
    if isinstance(data, np.ndarray):
        return data
    return np.array([float(v) for v in data])


def floatList2D(data):
    """2-D list of numbers / unit strings -> float ndarray."""
    if isinstance(data, np.ndarray):
        return data
    return np.array([[float(v) for v in row] for row in data])
