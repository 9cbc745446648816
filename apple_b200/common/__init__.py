"""Attribute names, Lame conversion and potential auto-naming.

Mirrors ``liblaf.apple.common`` (``common/attr_name.py:4-44``, ``common/_moduli.py:5-10``,
``common/_potential_name.py:6-16``).
"""

from __future__ import annotations

import enum
from collections import Counter


class AttrName(enum.StrEnum):
    """Names of mesh attributes; ``.vtk`` is the key used in ``point_data`` / ``cell_data``."""

    vtk: str
    ACTIVATION = enum.auto()
    DISPLACEMENT = enum.auto()
    E = enum.auto()
    FIXED_MASK = enum.auto()
    FIXED_VALUE = enum.auto()
    FORCE = enum.auto()
    FRACTION = enum.auto()
    GLOBAL_POINT_ID = enum.auto()
    LAMBDA = "lambda_"
    MASS_DENSITY = enum.auto()
    MU = enum.auto()
    NU = enum.auto()
    PRESTRAIN = enum.auto()


for _name in AttrName:
    _name.vtk = _name.value.title().replace("_", "")
AttrName.LAMBDA.vtk = "lambda"
AttrName.MU.vtk = "mu"
AttrName.NU.vtk = "nu"

ACTIVATION = AttrName.ACTIVATION
DISPLACEMENT = AttrName.DISPLACEMENT
E = AttrName.E
FIXED_MASK = AttrName.FIXED_MASK
FIXED_VALUE = AttrName.FIXED_VALUE
FORCE = AttrName.FORCE
FRACTION = AttrName.FRACTION
GLOBAL_POINT_ID = AttrName.GLOBAL_POINT_ID
LAMBDA = AttrName.LAMBDA
MASS_DENSITY = AttrName.MASS_DENSITY
MU = AttrName.MU
NU = AttrName.NU
PRESTRAIN = AttrName.PRESTRAIN


def lame_converter(E, nu):
    """(E, nu) -> (lambda, mu), ``common/_moduli.py:5-10``."""
    la = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))
    mu = E / (2.0 * (1.0 + nu))
    return la, mu


_counter: Counter[str] = Counter()


def default_potential_name(obj) -> str:
    """``ClassName<count>`` (``common/_potential_name.py:9-13``)."""
    cls_name = type(obj).__name__
    count = _counter[cls_name]
    _counter[cls_name] += 1
    return f"{cls_name}{count}"
