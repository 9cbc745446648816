"""``liblaf.peach.optim.pncg`` names used by the reference (``forward/_forward.py:24``)."""

from ._pncg import ConvergenceCriteria, LineSearch

__all__ = ["ConvergenceCriteria", "LineSearch"]
