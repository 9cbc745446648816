from __future__ import annotations

from collections.abc import Sequence

import torch

from apple_b200.common import default_potential_name


class WarpPotential:
    """Operator plugin interface, ``warp/model/_potential.py:9-32``.

    ``fun / grad / hess_diag / hess_prod / hess_quad`` ACCUMULATE into a caller-zeroed ``output``
    (CUDA tensors; ``(1,)`` for the scalars, ``(n_points, 3|4)`` for the fields) and return ``None``.
    """

    materials = None

    def __init__(self, *, name: str | None = None, requires_grad: Sequence[str] = ()):
        self.name = name if name is not None else default_potential_name(self)
        self.requires_grad = tuple(requires_grad)

    def fun(self, u: torch.Tensor, output: torch.Tensor) -> None:
        raise NotImplementedError

    def grad(self, u: torch.Tensor, output: torch.Tensor) -> None:
        raise NotImplementedError

    def hess_diag(self, u: torch.Tensor, output: torch.Tensor) -> None:
        raise NotImplementedError

    def hess_prod(self, u: torch.Tensor, p: torch.Tensor, output: torch.Tensor) -> None:
        raise NotImplementedError

    def hess_quad(self, u: torch.Tensor, p: torch.Tensor, output: torch.Tensor) -> None:
        raise NotImplementedError

    # fused form used by WarpModel.eval: any OR of the five operators in one pass
    def eval(self, ops: int, u, p, *, fun=None, quad=None, grad=None, diag=None, prod=None, scatter=None) -> None:
        from apple_b200 import _lib

        if ops & _lib.OP_FUN:
            self.fun(u, fun)
        if ops & _lib.OP_GRAD:
            self.grad(u, grad)
        if ops & _lib.OP_HESS_DIAG:
            self.hess_diag(u, diag)
        if ops & _lib.OP_HESS_PROD:
            self.hess_prod(u, p, prod)
        if ops & _lib.OP_HESS_QUAD:
            self.hess_quad(u, p, quad)
