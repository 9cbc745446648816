from types import SimpleNamespace

from apple_b200 import _lib

from ._base import WarpPotentialFem, get_lambda, get_mu


class StableNeoHookean(WarpPotentialFem):
    """Psi = mu/2 (I2 - 3) - mu (J - 1) + lambda/2 (J - 1)^2, ``warp/fem/_stable_neo_hookean.py:17-154``."""

    KIND = _lib.KIND_SNH
    MATERIAL_NAMES = ("lambda_", "mu")

    @classmethod
    def materials_from_region(cls, region, requires_grad):  # :148-154
        return SimpleNamespace(lambda_=get_lambda(region), mu=get_mu(region))
