"""Forward-simulation layer: mirror of ``liblaf.apple.forward``."""

from ._builder import ModelBuilder
from ._forward import Forward
from ._model import Model
from ._problem import ForwardProblem
from ._state import ModelState
from .dof_map import DofMap, DofMapBuilder

__all__ = ["DofMap", "DofMapBuilder", "Forward", "ForwardProblem", "Model", "ModelBuilder", "ModelState"]
