from ._builder import DofMapBuilder
from ._dof_map import DofMap

__all__ = ["DofMap", "DofMapBuilder"]
