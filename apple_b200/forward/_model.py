from __future__ import annotations

import torch

from apple_b200.warp.model import WarpModelAdapter

from ._state import ModelState
from .dof_map import DofMap


class Model:
    """Mirror of ``forward/_model.py:14-74``: elastic potentials plus an optional collision term."""

    State = ModelState

    def __init__(self, dof_map: DofMap, warp_model: WarpModelAdapter, collision=None):
        self.dof_map = dof_map
        self.warp_model = warp_model
        self.collision = collision

    @property
    def dim(self) -> int:
        return self.dof_map.dim

    @property
    def n_fixed(self) -> int:
        return self.dof_map.n_fixed

    @property
    def n_free(self) -> int:
        return self.dof_map.n_free

    @property
    def n_full(self) -> int:
        return self.dof_map.n_full

    @property
    def n_points(self) -> int:
        return self.dof_map.n_points

    def init(self) -> ModelState:
        free = torch.zeros(self.n_free, dtype=self.dof_map.fixed_values.dtype, device=self.dof_map.free_indices.device)
        return ModelState(u=self.dof_map.to_full(free))

    def fun(self, u):
        output = self.warp_model.fun(u)
        if self.collision is not None:
            output = output + self.collision.fun(u)
        return output

    def grad(self, u):
        output = self.warp_model.grad(u)
        if self.collision is not None:
            output = output + self.collision.grad(u)
        return output

    def hess_diag(self, u):
        output = self.warp_model.hess_diag(u)
        if self.collision is not None:
            output = output + self.collision.hess_diag(u)
        return output

    def hess_prod(self, u, p):
        output = self.warp_model.hess_prod(u, p)
        if self.collision is not None:
            output = output + self.collision.hess_prod(u, p)
        return output

    def hess_quad(self, u, p):
        output = self.warp_model.hess_quad(u, p)
        if self.collision is not None:
            output = output + self.collision.hess_quad(u, p)
        return output
