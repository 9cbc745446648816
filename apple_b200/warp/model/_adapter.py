from __future__ import annotations

import torch

from apple_b200 import _lib

from ._model import WarpModel


class WarpModelAdapter:
    """Pure-function view of a ``WarpModel``, ``warp/model/_adapter.py:16-140``.

    The reference wraps five JAX-FFI callables; here the same five functions allocate their output
    on ``u``'s device and call straight into the CUDA kernels on the current torch stream.  dtype is
    taken from ``u`` like the reference's dtype-generic factories (``_adapter.py:46-47``)."""

    def __init__(self, model: WarpModel, n_points: int):
        self.__wrapped__ = model
        self.n_points = int(n_points)

    def _scalar(self, u: torch.Tensor) -> torch.Tensor:
        return torch.empty(1, dtype=u.dtype, device=u.device)

    def _field(self, u: torch.Tensor) -> torch.Tensor:
        return torch.empty((self.n_points, 3), dtype=u.dtype, device=u.device)

    def fun(self, u: torch.Tensor) -> torch.Tensor:
        u = u.contiguous()
        output = self._scalar(u)
        self.__wrapped__.fun(u, output)
        return output[0]

    def grad(self, u: torch.Tensor) -> torch.Tensor:
        u = u.contiguous()
        output = self._field(u)
        self.__wrapped__.grad(u, output)
        return output

    def hess_diag(self, u: torch.Tensor) -> torch.Tensor:
        u = u.contiguous()
        output = self._field(u)
        self.__wrapped__.hess_diag(u, output)
        return output

    def hess_prod(self, u: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
        u, p = u.contiguous(), p.contiguous()
        output = self._field(u)
        self.__wrapped__.hess_prod(u, p, output)
        return output

    def hess_quad(self, u: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
        u, p = u.contiguous(), p.contiguous()
        output = self._scalar(u)
        self.__wrapped__.hess_quad(u, p, output)
        return output[0]

    # ---- fused forms (one pass over the elements per potential) ----
    def fun_grad_hess_prod(self, u: torch.Tensor, p: torch.Tensor, *, scatter=None):
        """(energy, gradient, Hessian-vector product): the fused evaluation of the headline metric."""
        u, p = u.contiguous(), p.contiguous()
        fun, grad, prod = self._scalar(u), self._field(u), self._field(u)
        self.__wrapped__.eval(
            _lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_PROD, u, p, fun=fun, grad=grad, prod=prod, scatter=scatter
        )
        return fun[0], grad, prod

    def fun_grad_hess_diag(self, u: torch.Tensor, *, scatter=None):
        """(energy, gradient, Hessian diagonal): PNCG's pass A."""
        u = u.contiguous()
        fun, grad, diag = self._scalar(u), self._field(u), self._field(u)
        self.__wrapped__.eval(
            _lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_DIAG, u, None, fun=fun, grad=grad, diag=diag, scatter=scatter
        )
        return fun[0], grad, diag
