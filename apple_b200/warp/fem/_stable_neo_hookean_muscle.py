from types import SimpleNamespace

from apple_b200 import _lib

from ._base import WarpPotentialFem, get_activation, get_lambda, get_mu


class StableNeoHookeanMuscle(WarpPotentialFem):
    """Stable Neo-Hookean on G = F A with a per-cell symmetric activation A,
    ``warp/fem/_stable_neo_hookean_muscle.py:18-171``."""

    KIND = _lib.KIND_SNH_MUSCLE
    MATERIAL_NAMES = ("activation", "lambda_", "mu")

    @classmethod
    def materials_from_region(cls, region, requires_grad):  # :163-171
        return SimpleNamespace(activation=get_activation(region), lambda_=get_lambda(region), mu=get_mu(region))
