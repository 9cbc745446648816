from types import SimpleNamespace

from apple_b200 import _lib

from ._base import WarpPotentialFem, get_mu


class Arap(WarpPotentialFem):
    """Psi = mu/2 |F - R|^2, ``warp/fem/_arap.py:17-125``.

    ``hess_prod`` is the mathematically correct Hessian-vector product (with the reference's clamped
    twist eigenvalues); the reference swaps two arguments there (``_arap.py:55-56``), see DESIGN.md."""

    KIND = _lib.KIND_ARAP
    MATERIAL_NAMES = ("mu",)

    @classmethod
    def materials_from_region(cls, region, requires_grad):  # :120-125
        return SimpleNamespace(mu=get_mu(region))
