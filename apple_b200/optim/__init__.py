"""Optimizer layer: the part of ``liblaf.peach.optim`` the reference uses (``Problem``, ``Optimizer``,
``PNCG``, ``Result``).  peach itself is an external, un-vendored dependency of the reference
(``pyproject.toml:39,172``); its call sites are ``forward/_forward.py:22-30,52-56`` and
``forward/_problem.py:18-59``."""

from ._base import Optimizer, Problem, Result, Solution
from ._pcg import PcgInfo, adjoint_solve, pcg
from ._pncg import PNCG
from . import pncg

__all__ = ["PNCG", "Optimizer", "PcgInfo", "Problem", "Result", "Solution", "adjoint_solve", "pcg", "pncg"]
