"""Preconditioned nonlinear conjugate gradient (PNCG).

The reference's optimizer is ``liblaf.peach.optim.PNCG`` -- external, un-vendored, un-pinned
(``pyproject.toml:39,172``) -- so its internals are NOT pinned.  What is implemented are the
recurrences of the reference's own PNCG-like benchmark
(``benches/bench_pncg_branching_backends.py:254-329,407-410,606-610,663-679``) and Shen et al. 2024,
exposing the state fields the reference's drivers read
(``exp/2026/05/06/toy/src/20-...collision.py:207-233``): ``line_search_state.{alpha,f0,f_alpha,step,ok}``,
``convergence_state.{grad_norm,grad_norm_first}``, ``hess_damping_state.{factor,hess_diag_mean}``,
``direction``, ``n_steps``.

Two execution paths, same recurrences:

* FUSED (the product path): taken for a ``ForwardProblem`` without collision.  The whole iteration
  runs inside ``libapple_b200.so`` (``apple_b200/csrc/pncg.cu``): fixed DOFs are masked rather than
  gathered/scattered, every scalar stays on the device, the line-search trial point is formed inside
  the element kernel's gather, and the host only reads the scalars every ``check_every`` iterations.
* GENERIC: any other ``Problem`` (e.g. with a collision term evaluated by a CPU callback).  The
  operators are the problem's own callbacks; the handful of vector operations between them are
  torch ops on the problem's device.
"""

from __future__ import annotations

import ctypes
import math
import time
from dataclasses import dataclass
from types import SimpleNamespace

import torch

from apple_b200 import _lib

from ._base import Optimizer, Result, Solution


@dataclass
class ConvergenceCriteria:
    """Termination rules.  ``max_steps`` is the reference's default budget (``forward/_forward.py:26-27``);
    the gradient-norm thresholds carry the names the reference's experiments set
    (``exp/2026/01/28/smas/src/31-inverse-activation-stable-neo-hookean.py:212-218``)."""

    max_steps: int = 1500
    target_relative_gradient_norm: float = 1.0e-7      # -> PRIMARY_SUCCESS
    acceptable_relative_gradient_norm: float = 1.0e-3  # -> SECONDARY_SUCCESS when the run stops otherwise
    absolute_gradient_norm: float = 0.0
    max_failed_line_searches: int = 5                  # consecutive; then the run stops (stagnation)


@dataclass
class LineSearch:
    """Newton-step initial length + Armijo backtracking (bench ``:606-610``, ``:413-456``)."""

    overstep: float = 1.0
    armijo: float = 1.0e-4
    max_steps: int = 8  # halvings


class PNCG(Optimizer):
    def __init__(self, criteria: ConvergenceCriteria | None = None, line_search: LineSearch | None = None, *,
                 fused: bool = True, check_every: int | None = None, use_graph: bool | int = 2,
                 scatter: int | None = None, preconditioner: str = "jacobi", psd: bool = False):
        """``preconditioner``: ``"jacobi"`` -- the reference's scalar rule on the clamped Hessian diagonal (bench
        ``:407-410``; the parity default) -- or ``"block"``, the opt-in 3x3 block Jacobi on the vertex blocks of the
        Hessian.  ``psd=True`` (opt-in): passes A and B use the eigenvalue-clamped element Hessians.  Both are fused-path
        features (``apl_pncg_set_block_jacobi``)."""
        if preconditioner not in ("jacobi", "block"):
            raise ValueError("preconditioner must be 'jacobi' or 'block'")
        self.preconditioner = preconditioner
        self.psd = bool(psd)
        self.criteria = criteria if criteria is not None else ConvergenceCriteria()
        self.line_search = line_search if line_search is not None else LineSearch()
        self.fused = fused
        self.check_every = check_every
        self.use_graph = use_graph
        self.scatter = scatter

    # the reference's experiments call the criteria ``convergence``
    @property
    def convergence(self) -> ConvergenceCriteria:
        return self.criteria

    @convergence.setter
    def convergence(self, value: ConvergenceCriteria) -> None:
        self.criteria = value

    # ------------------------------------------------------------------ protocol
    def init(self, problem, state, free):
        if self.fused and _fused_supported(problem):
            return _FusedState(self, problem, state, free)
        if self.preconditioner != "jacobi" or self.psd:
            raise NotImplementedError("block Jacobi / PSD projection exist on the fused PNCG path only")
        return _GenericState(self, problem, state, free)

    def step(self, problem, state, opt_state):
        state = opt_state.step(problem, state, 1)
        return state, opt_state

    def terminate(self, problem, state, opt_state):
        return opt_state.terminate()

    def postprocess(self, problem, state, opt_state, result) -> Solution:
        return Solution(
            result=result,
            state=state,
            params=opt_state.free(problem, state),
            stats={
                "n_steps": opt_state.n_steps,
                "n_accepted": opt_state.n_accepted,
                "fun": opt_state.line_search_state.f_alpha,
                "relative_grad_norm": opt_state.relative_grad_norm,
                "time": time.perf_counter() - opt_state.t_start,
                "fused": isinstance(opt_state, _FusedState),
            },
        )

    def minimize(self, problem, state, free):
        opt_state = self.init(problem, state, free)
        chunk = self.check_every
        if chunk is None:
            chunk = 1 if free.numel() < 50_000 else 16
        while True:
            state = opt_state.step(problem, state, chunk)
            done, result = opt_state.terminate()
            if done:
                break
        return self.postprocess(problem, state, opt_state, result), state


def _fused_supported(problem) -> bool:
    from apple_b200.forward import ForwardProblem
    from apple_b200.warp.fem import WarpPotentialFem
    from apple_b200.warp.potential import ExternalForce

    if not isinstance(problem, ForwardProblem) or problem.model.collision is not None:
        return False
    pots = list(problem.model.warp_model.__wrapped__.potentials.values())
    return bool(pots) and all(isinstance(p, (WarpPotentialFem, ExternalForce)) for p in pots)


def _classify(done_code: float, rel: float, criteria: ConvergenceCriteria) -> Result:
    if done_code == 1.0:
        return Result.PRIMARY_SUCCESS
    if done_code == 4.0:
        return Result.NAN_ENCOUNTERED
    if rel <= criteria.acceptable_relative_gradient_norm:
        return Result.SECONDARY_SUCCESS
    if done_code == 2.0:
        return Result.MAX_STEPS_REACHED
    if done_code == 3.0:
        return Result.STAGNATION
    return Result.UNKNOWN_ERROR


class _StateBase:
    n_steps = 0
    n_accepted = 0

    def __init__(self):
        self.t_start = time.perf_counter()
        self.line_search_state = SimpleNamespace(alpha=0.0, f0=math.nan, f_alpha=math.nan, step=0, ok=True)
        self.convergence_state = SimpleNamespace(grad_norm=math.nan, grad_norm_first=math.nan)
        self.hess_damping_state = SimpleNamespace(factor=0.0, hess_diag_mean=math.nan)
        self.done_code = 0.0

    @property
    def relative_grad_norm(self) -> float:
        c = self.convergence_state
        return c.grad_norm / c.grad_norm_first if c.grad_norm_first > 0 else 0.0


class _FusedState(_StateBase):
    """Device-resident PNCG state driving ``apl_pncg_*`` (``apple_b200/csrc/pncg.cu``)."""

    def __init__(self, opt: PNCG, problem, state, free):
        super().__init__()
        from apple_b200.warp.fem import WarpPotentialFem

        model = problem.model
        dm = model.dof_map
        u = state.u
        self.dtype, self.device = u.dtype, u.device
        self.n_points = n = model.n_points
        self.criteria = opt.criteria
        new = lambda: torch.zeros((n, 4), dtype=self.dtype, device=self.device)  # noqa: E731
        self.x = new()
        self.x[:, :3] = dm.to_full(free.to(self.dtype))
        self.p = [new(), new()]
        self.g = [new(), new()]
        self.d = [new(), new()]
        mask = torch.zeros((n, 4), dtype=torch.uint8, device=self.device)
        mask[:, :3] = dm.free_mask().to(torch.uint8) * 3  # free | counted
        self.mask = mask
        self.scal = torch.zeros(_lib.PNCG_NSCAL, dtype=torch.float64, device=self.device)
        self._scal_host = torch.zeros(_lib.PNCG_NSCAL, dtype=torch.float64).pin_memory()
        handle = ctypes.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        L = _lib.lib()
        _lib.check(L.apl_pncg_create(
            _lib.dtype_code(self.dtype), n, dev_index, _lib.dev_ptr(self.x), _lib.dev_ptr(self.p[0]),
            _lib.dev_ptr(self.p[1]), _lib.dev_ptr(self.g[0]), _lib.dev_ptr(self.g[1]), _lib.dev_ptr(self.d[0]),
            _lib.dev_ptr(self.d[1]), _lib.dev_ptr(self.mask), _lib.dev_ptr(self.scal), ctypes.byref(handle)))
        self._handle = handle
        self._keep = []
        for pot in model.warp_model.__wrapped__.potentials.values():
            if isinstance(pot, WarpPotentialFem):
                if pot.dtype != self.dtype:
                    raise TypeError(f"potential {pot.name} is {pot.dtype}, state is {self.dtype}")
                _lib.check(L.apl_pncg_add_fem(handle, pot._handle))
            else:
                if pot.dtype != self.dtype:
                    raise TypeError(f"potential {pot.name} is {pot.dtype}, state is {self.dtype}")
                if getattr(pot, "_max_index", -1) >= n:
                    raise IndexError(f"potential {pot.name} loads vertex {pot._max_index}, the model has {n} points")
                _lib.check(L.apl_pncg_add_ext_force(handle, pot.indices.shape[0], _lib.dev_ptr(pot.materials.force),
                                                    _lib.dev_ptr(pot.indices)))
            self._keep.append(pot)
        c, ls = opt.criteria, opt.line_search
        from apple_b200 import config

        scatter = opt.scatter if opt.scatter is not None else config.scatter
        _lib.check(L.apl_pncg_set_params(
            handle, float(c.max_steps), float(c.target_relative_gradient_norm), float(c.absolute_gradient_norm),
            float(c.max_failed_line_searches), float(ls.overstep), 1.0, float(ls.armijo), int(ls.max_steps),
            int(scatter), int(opt.use_graph) if opt.use_graph in (0, 1, 2) else int(bool(opt.use_graph))))
        self.o = [new(), new()] if opt.preconditioner == "block" else [None, None]
        if opt.preconditioner == "block" or opt.psd:
            _lib.check(L.apl_pncg_set_block_jacobi(handle, _lib.dev_ptr(self.o[0]), _lib.dev_ptr(self.o[1]), int(opt.psd)))
        with torch.cuda.device(self.device):
            _lib.check(L.apl_pncg_phase(handle, _lib.PHASE_INIT, 0, _lib.stream_ptr(self.device)))

    def __del__(self):
        handle = getattr(self, "_handle", None)
        if handle is not None and handle.value:
            try:
                _lib.lib().apl_pncg_destroy(handle)
            except Exception:
                pass
            self._handle = None

    def _read_scalars(self):
        self._scal_host.copy_(self.scal, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._scal_host

    @property
    def direction(self) -> torch.Tensor:
        """Free-DOF view of the last search direction."""
        cur = int(self._scal_host[_lib.S_K].item()) % 2
        last = self.p[1 - cur] if self.n_steps > 0 else self.p[cur]
        return last[:, :3].reshape(-1)[self._free_indices()]

    def _free_indices(self):
        return (self.mask[:, :3].reshape(-1) & 1).nonzero().reshape(-1)

    def step(self, problem, state, n_iters: int):
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().apl_pncg_iterate(self._handle, int(n_iters), _lib.stream_ptr(self.device)))
        s = self._read_scalars()
        self.n_steps = int(s[_lib.S_K].item())
        self.n_accepted = int(s[_lib.S_N_ACCEPTED].item())
        self.done_code = float(s[_lib.S_DONE].item())
        ls = self.line_search_state
        ls.alpha = float(s[_lib.S_ALPHA].item())
        ls.f_alpha = float(s[_lib.S_F].item())
        ls.ok = bool(s[_lib.S_ACCEPTED].item())
        ls.f0 = float(s[_lib.S_F_PREV].item()) if ls.ok else ls.f_alpha
        ls.step = int(s[_lib.S_LS_STEPS].item())
        cs = self.convergence_state
        cs.grad_norm = math.sqrt(max(float(s[_lib.S_GNORM2].item()), 0.0))
        cs.grad_norm_first = math.sqrt(max(float(s[_lib.S_GNORM2_FIRST].item()), 0.0))
        self.hess_damping_state.hess_diag_mean = float(s[_lib.S_DIAG_MEAN].item())
        import dataclasses

        return dataclasses.replace(state, u=self.x[:, :3].contiguous())

    def terminate(self):
        if self.done_code != 0.0:
            return True, _classify(self.done_code, self.relative_grad_norm, self.criteria)
        return False, Result.UNKNOWN_ERROR

    def free(self, problem, state):
        return problem.model.dof_map.to_free(state.u)


class _GenericState(_StateBase):
    """Host-driven PNCG over an arbitrary ``Problem``: one host decision per iteration."""

    def __init__(self, opt: PNCG, problem, state, free):
        super().__init__()
        self.opt = opt
        self.criteria = opt.criteria
        self.x = free.clone()
        self.g_prev = torch.zeros_like(free)
        self.p_prev = torch.zeros_like(free)
        self.direction = torch.zeros_like(free)
        self.fails = 0
        self.f = None

    def _one(self, problem, state):
        c, ls = self.criteria, self.opt.line_search
        state = problem.before_trial(state, self.x)
        f = float(problem.fun(state)) if self.f is None else self.f
        g = problem.grad(state)
        gnorm = float(torch.linalg.vector_norm(g))
        cs = self.convergence_state
        if self.n_steps == 0:
            cs.grad_norm_first = gnorm
        cs.grad_norm = gnorm
        if not math.isfinite(gnorm):
            self.done_code = 4.0
        elif gnorm <= c.absolute_gradient_norm or gnorm <= c.target_relative_gradient_norm * cs.grad_norm_first:
            self.done_code = 1.0
        elif self.n_steps >= c.max_steps:
            self.done_code = 2.0
        elif self.fails >= c.max_failed_line_searches:
            self.done_code = 3.0
        if self.done_code:
            self.f = f
            self.line_search_state.f_alpha = f
            return state
        d = problem.hess_diag(state).abs()
        pos = d > 0
        mean = d[pos].mean() if bool(pos.any()) else torch.ones((), dtype=d.dtype, device=d.device)
        P = 1.0 / torch.where(pos, d, mean)
        self.hess_damping_state.hess_diag_mean = float(mean)
        beta = 0.0
        if self.n_steps > 0:
            y = g - self.g_prev
            yp = float(torch.dot(y, self.p_prev))
            if abs(yp) > 1e-12:
                Py = P * y
                beta = float(torch.dot(g, Py)) / yp - (float(torch.dot(y, Py)) / yp) * (float(torch.dot(self.p_prev, g)) / yp)
            else:
                beta = math.inf
            if not math.isfinite(beta) or abs(beta) > 10.0:
                beta = 0.0
        steepest = -P * g
        p = steepest + beta * self.p_prev
        gp = float(torch.dot(g, p))
        if not (math.isfinite(gp) and gp < 0.0):
            p, beta = steepest, 0.0
            gp = float(torch.dot(g, p))
        pHp = float(problem.hess_quad(state, p))
        alpha = -gp / pHp if pHp != 0 else math.copysign(math.inf, -gp) if gp != 0 else math.nan
        if math.isnan(alpha) or alpha == -math.inf:
            alpha = 0.0
        elif alpha == math.inf:
            alpha = 1.0
        if not (alpha > 0.0 and math.isfinite(alpha)):
            alpha = 1.0
        alpha = min(alpha * ls.overstep, float(problem.max_step_size(state, p)))
        tries, accepted = 0, False
        while True:
            x_trial = self.x + alpha * p
            trial_state = problem.before_trial(state, x_trial)
            f_trial = float(problem.fun(trial_state))
            accepted = math.isfinite(f_trial) and f_trial <= f + ls.armijo * alpha * gp
            if accepted or tries >= ls.max_steps or not alpha > 0.0:
                break
            alpha *= 0.5
            tries += 1
        lss = self.line_search_state
        lss.alpha, lss.f0, lss.step, lss.ok = alpha, f, tries, accepted
        if accepted:
            self.x, self.f, state = x_trial, f_trial, trial_state
            self.n_accepted += 1
            self.fails = 0
        else:
            self.f = f
            self.fails += 1
        lss.f_alpha = self.f
        self.g_prev, self.p_prev, self.direction = g, p, p
        self.n_steps += 1
        return state

    def step(self, problem, state, n_iters: int):
        for _ in range(n_iters):
            state = self._one(problem, state)
            if self.done_code:
                break
        return state

    def terminate(self):
        if self.done_code != 0.0:
            return True, _classify(self.done_code, self.relative_grad_norm, self.criteria)
        return False, Result.UNKNOWN_ERROR

    def free(self, problem, state):
        return self.x
