from ._adapter import WarpModelAdapter
from ._model import WarpModel
from ._potential import WarpPotential

__all__ = ["WarpModel", "WarpModelAdapter", "WarpPotential"]
