from __future__ import annotations

import dataclasses

import torch

from apple_b200.optim import Problem

from ._model import Model
from ._state import ModelState


class ForwardProblem(Problem):
    """Mirror of ``forward/_problem.py:18-59``: the optimisation problem PNCG sees (free DOFs)."""

    State = ModelState

    def __init__(self, model: Model):
        self.model = model

    def before_trial(self, state: ModelState, u: torch.Tensor) -> ModelState:  # :24-27
        return dataclasses.replace(state, u=self.model.dof_map.to_full(u))

    def max_step_size(self, state: ModelState, p: torch.Tensor):  # :29-34
        if self.model.collision is None:
            return torch.ones((), dtype=p.dtype, device=p.device)
        p_full = self.model.dof_map.to_full_grad(p)
        return self.model.collision.max_step_size(state.u, p_full)

    def fun(self, state: ModelState):  # :36-38
        return self.model.fun(state.u)

    def grad(self, state: ModelState):  # :40-43
        return self.model.dof_map.to_free_grad(self.model.grad(state.u))

    def hess_diag(self, state: ModelState):  # :45-48
        return self.model.dof_map.to_free_hess_diag(self.model.hess_diag(state.u))

    def hess_prod(self, state: ModelState, p):  # :50-54
        p_full = self.model.dof_map.to_full_grad(p)
        return self.model.dof_map.to_free_grad(self.model.hess_prod(state.u, p_full))

    def hess_quad(self, state: ModelState, p):  # :56-59
        p_full = self.model.dof_map.to_full_grad(p)
        return self.model.hess_quad(state.u, p_full)
