from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Any


class Result(enum.Enum):
    """Termination classes.  ``PRIMARY_SUCCESS`` and ``UNKNOWN_ERROR`` are the names the reference
    uses (``tests/forward/test_static_simulation.py:85``,
    ``exp/2026/05/06/toy/src/20-ext-force-stable-neo-hookean-collision.py:242``)."""

    PRIMARY_SUCCESS = enum.auto()    # target criterion met
    SECONDARY_SUCCESS = enum.auto()  # acceptable criterion met when the budget ran out / progress stalled
    MAX_STEPS_REACHED = enum.auto()
    STAGNATION = enum.auto()
    NAN_ENCOUNTERED = enum.auto()
    UNKNOWN_ERROR = enum.auto()


@dataclass
class Solution:
    result: Result
    state: Any
    params: Any
    stats: dict = field(default_factory=dict)

    @property
    def success(self) -> bool:
        return self.result in (Result.PRIMARY_SUCCESS, Result.SECONDARY_SUCCESS)


class Problem:
    """The protocol ``ForwardProblem`` implements (``forward/_problem.py:24-59``).  Vectors are
    free-DOF tensors; ``state`` is an opaque model state."""

    def before_trial(self, state, x):
        raise NotImplementedError

    def max_step_size(self, state, p):
        raise NotImplementedError

    def fun(self, state):
        raise NotImplementedError

    def grad(self, state):
        raise NotImplementedError

    def hess_diag(self, state):
        raise NotImplementedError

    def hess_prod(self, state, p):
        raise NotImplementedError

    def hess_quad(self, state, p):
        raise NotImplementedError


class Optimizer:
    """Stepping protocol used by the reference's experiment drivers
    (``exp/2026/05/06/toy/src/20-ext-force-stable-neo-hookean-collision.py:236-253``)."""

    Solution = Solution

    def init(self, problem, state, free):
        raise NotImplementedError

    def step(self, problem, state, opt_state):
        raise NotImplementedError

    def terminate(self, problem, state, opt_state):
        raise NotImplementedError

    def postprocess(self, problem, state, opt_state, result):
        raise NotImplementedError

    def minimize(self, problem, state, free):
        opt_state = self.init(problem, state, free)
        result = Result.UNKNOWN_ERROR
        while True:
            state, opt_state = self.step(problem, state, opt_state)
            done, result = self.terminate(problem, state, opt_state)
            if done:
                break
        return self.postprocess(problem, state, opt_state, result), state
