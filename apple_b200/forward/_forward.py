from __future__ import annotations

import functools

import torch

from apple_b200.optim import PNCG, Optimizer
from apple_b200.optim.pncg import ConvergenceCriteria, LineSearch

from ._model import Model
from ._problem import ForwardProblem
from ._state import ModelState


class Forward:
    """Mirror of ``forward/_forward.py:18-56``: one static solve per ``step()``."""

    def __init__(self, model: Model, optimizer: Optimizer | None = None, state: ModelState | None = None):
        self.model = model
        self.optimizer = optimizer if optimizer is not None else self._default_optimizer()
        self.state = state if state is not None else self._default_state()

    def _default_optimizer(self) -> Optimizer:  # :22-30
        return PNCG(criteria=ConvergenceCriteria(max_steps=1500), line_search=LineSearch())

    def _default_state(self) -> ModelState:  # :32-35
        dm = self.model.dof_map
        u_free = torch.zeros(self.model.n_free, dtype=dm.fixed_values.dtype, device=dm.free_indices.device)
        return ModelState(u=dm.to_full(u_free))

    @property
    def free(self) -> torch.Tensor:
        return self.model.dof_map.to_free(self.state.u)

    @functools.cached_property
    def problem(self) -> ForwardProblem:
        return ForwardProblem(model=self.model)

    def step(self):  # :52-56
        solution, self.state = self.optimizer.minimize(self.problem, self.state, self.free)
        return solution
