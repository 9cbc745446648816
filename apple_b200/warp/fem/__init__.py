from ._arap import Arap
from ._base import WarpPotentialFem
from ._stable_neo_hookean import StableNeoHookean
from ._stable_neo_hookean_muscle import StableNeoHookeanMuscle

__all__ = ["Arap", "StableNeoHookean", "StableNeoHookeanMuscle", "WarpPotentialFem"]
