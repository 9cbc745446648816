from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from apple_b200 import _lib, config
from apple_b200.common import FORCE, GLOBAL_POINT_ID
from apple_b200.warp.model import WarpPotential


class ExternalForce(WarpPotential):
    """W = -sum_k f_k . u[idx_k]; grad[idx_k] -= f_k; zero Hessian.
    Mirror of ``warp/potential/_ext_force.py:42-90``."""

    def __init__(self, indices, force, *, dtype=None, device=None, name=None):
        super().__init__(name=name)
        self.dtype = dtype or config.default_dtype
        self.device = torch.device(device if device is not None else config.default_device())
        if self.device.type != "cuda":
            raise _lib.NativeError("apple_b200 potentials live on a CUDA device (there is no CPU path)")
        self.indices = torch.as_tensor(np.asarray(indices, dtype=np.int32), device=self.device).contiguous()
        self.materials = SimpleNamespace(
            force=torch.as_tensor(np.asarray(force, dtype=_lib.np_dtype(self.dtype)), device=self.device).contiguous()
        )
        if self.materials.force.shape != (self.indices.shape[0], 3):
            raise ValueError("force must have shape (len(indices), 3)")
        idx = np.asarray(indices)
        if idx.size and int(idx.min()) < 0:
            raise IndexError("ExternalForce: negative vertex index")
        self._max_index = int(idx.max()) if idx.size else -1

    @classmethod
    def from_pyvista(cls, obj, **kwargs):  # :55-61
        force = np.asarray(obj.point_data[FORCE.vtk])
        indices = np.asarray(obj.point_data[GLOBAL_POINT_ID.vtk])
        return cls(indices, force, **kwargs)

    def eval(self, ops: int, u, p=None, *, fun=None, quad=None, grad=None, diag=None, prod=None, scatter=None,
             part: int = 0) -> None:
        ops &= _lib.OP_FUN | _lib.OP_GRAD  # Hessian operators are no-ops (:80-90)
        if not ops or part == _lib.PART_INTERIOR:   # a split evaluation applies the loads with the boundary part
            return
        # the kernel reinterprets raw pointers as `self.dtype`: a field of another dtype / layout must not get through
        n_points = int(u.shape[0]) if u.dim() == 2 else -1
        ld_in = _lib.field_ld(u, n_points, self.dtype, "u")
        ld_out = _lib.field_ld(grad, n_points, self.dtype, "grad") if (ops & _lib.OP_GRAD) else 3
        if (ops & _lib.OP_FUN) and (fun is None or fun.dtype != self.dtype or fun.numel() < 1):
            raise TypeError(f"fun: expected a ({self.dtype}) tensor with one element")
        if self._max_index >= n_points:
            raise IndexError(f"ExternalForce loads vertex {self._max_index}, the field has {n_points} points")
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.lib().apl_ext_force_eval(
                    _lib.dtype_code(self.dtype), ops, self.indices.shape[0], _lib.dev_ptr(self.materials.force),
                    _lib.dev_ptr(self.indices), _lib.dev_ptr(u), ld_in, _lib.dev_ptr(fun) if ops & _lib.OP_FUN else None,
                    _lib.dev_ptr(grad) if ops & _lib.OP_GRAD else None, ld_out, _lib.stream_ptr(self.device),
                )
            )

    def fun(self, u, output) -> None:  # :63-70
        self.eval(_lib.OP_FUN, u, fun=output)

    def grad(self, u, output) -> None:  # :72-78
        self.eval(_lib.OP_GRAD, u, grad=output)

    def hess_diag(self, u, output) -> None:  # :80-82
        pass

    def hess_prod(self, u, p, output) -> None:  # :84-86
        pass

    def hess_quad(self, u, p, output) -> None:  # :88-90
        pass
