from ._arap import Arap
from ._base import WarpPotentialFem
from ._fused import FusedSnhArap, fuse_potentials
from ._stable_neo_hookean import StableNeoHookean
from ._stable_neo_hookean_muscle import StableNeoHookeanMuscle

__all__ = ["Arap", "FusedSnhArap", "StableNeoHookean", "StableNeoHookeanMuscle", "WarpPotentialFem", "fuse_potentials"]
