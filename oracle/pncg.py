"""Oracle: the PNCG iteration (test infrastructure).  PARITY UNPINNED at the iteration level.

The optimizer the reference uses, ``liblaf.peach.optim.PNCG``, is an un-vendored, un-pinned
dependency (``pyproject.toml:39,172``); its source is not under ``/root/reference``.  What is
restated here is the only in-tree statement of the recurrences, the reference's "PNCG-like"
benchmark ``benches/bench_pncg_branching_backends.py``:

* preconditioner fix-up     ``:407-410``  (|d|, non-positive -> mean of positive, reciprocal)
* Dai-Kou beta              ``:663-679``
* beta reset rules          ``:288-289``
* direction + descent guard ``:290-303``
* initial step              ``:606-610``  (alpha = -(g.p)/pHp, sanitised, times overstep)
* Armijo backtracking       ``:413-456``
* bookkeeping               ``:320-324``

together with the problem glue of ``forward/_problem.py:24-59`` (free <-> full maps,
``max_step_size == 1`` without collision) and the published method (Shen et al., "Preconditioned
Nonlinear Conjugate Gradient Method for Real-time Interior-point Hyperelasticity", 2024).
Validation is at the convergence level (the reference's known-answer test).
"""

from __future__ import annotations

import numpy as np


class ForwardProblem:
    """``ForwardProblem`` (forward/_problem.py:18-59) over an ``oracle.fem.Model`` and an
    ``oracle.region.DofMap``; vectors are free-DOF numpy arrays."""

    def __init__(self, model, dof_map):
        self.model = model
        self.dof_map = dof_map

    def before_trial(self, x):  # :24-27
        return self.dof_map.to_full(x)

    def max_step_size(self, x, p):  # :29-34 (collision is None)
        return 1.0

    def fun(self, x):  # :36-38
        return self.model.fun(self.before_trial(x))

    def grad(self, x):  # :40-43
        return self.dof_map.to_free(self.model.grad(self.before_trial(x)))

    def hess_diag(self, x):  # :45-48
        return self.dof_map.to_free(self.model.hess_diag(self.before_trial(x)))

    def hess_prod(self, x, p):  # :50-54
        return self.dof_map.to_free(
            self.model.hess_prod(self.before_trial(x), self.dof_map.to_full_grad(p))
        )

    def hess_quad(self, x, p):  # :56-59
        return self.model.hess_quad(self.before_trial(x), self.dof_map.to_full_grad(p))

    def hess_quad_psd(self, x, p):
        """opt-in superset: the quadratic form of the PSD-projected element Hessians (``oracle/hessian.py``)"""
        from . import hessian as ohess

        u, pf = self.before_trial(x), self.dof_map.to_full_grad(p)
        return sum(ohess.hess_quad(pot, u, pf, psd=True) for pot in self.model.potentials if hasattr(pot, "cells"))


def make_preconditioner(hess_diag):
    """bench :407-410."""
    d = np.abs(hess_diag)
    pos = d > 0.0
    mean = d[pos].mean() if pos.any() else d.dtype.type(1.0)
    d = np.where(pos, d, mean)
    return 1.0 / d


class _Diagonal:
    """The reference's preconditioner as an operator (``P * v``)."""

    def __init__(self, P):
        self.P = P

    def __call__(self, v):
        return self.P * v


class BlockJacobi:
    """OPT-IN superset (not in the reference; BASELINE.json north star): inverse of the 3x3 vertex blocks of the
    assembled Hessian restricted to each vertex's free components; a block that is not positive definite falls back to
    the reference's scalar rule (bench ``:407-410``) on that vertex.  Blocks come from ``oracle/hessian.py``; their
    diagonal is the model's (clamped) ``hess_diag``."""

    def __init__(self, problem, x, psd=False):
        from . import hessian as ohess

        dm, model = problem.dof_map, problem.model
        u = problem.before_trial(x)
        V = dm.n_points
        B = np.zeros((V, 3, 3))
        for pot in model.potentials:
            if hasattr(pot, "cells"):
                B += ohess.vertex_blocks(pot, u, V, psd=psd)
        d = model.hess_diag(u) if not psd else np.stack([B[:, 0, 0], B[:, 1, 1], B[:, 2, 2]], 1)
        free = np.zeros(V * 3, bool)
        free[dm.free_indices] = True
        free = free.reshape(V, 3)
        A = B.copy()
        for c in range(3):
            A[:, c, c] = d[:, c]
        A = np.where(free[:, :, None] & free[:, None, :], A, 0.0)
        for c in range(3):
            A[~free[:, c], c, c] = 1.0
        m2 = A[:, 0, 0] * A[:, 1, 1] - A[:, 0, 1] ** 2
        det = np.linalg.det(A)
        self.spd = (A[:, 0, 0] > 0) & (m2 > 0) & (det > 0)
        self.inv = np.zeros_like(A)
        self.inv[self.spd] = np.linalg.inv(A[self.spd])
        dabs = np.abs(d)
        pos = (dabs > 0) & free
        mean = dabs[pos].mean() if pos.any() else 1.0
        self.scalar = 1.0 / np.where(dabs > 0, dabs, mean)
        self.dm = dm

    def __call__(self, v):
        full = self.dm.to_full_grad(v)
        out = np.where(self.spd[:, None], np.einsum("vij,vj->vi", self.inv, full), self.scalar * full)
        return self.dm.to_free(out)


def dai_kou_beta(g, g_prev, p_prev, P):
    """bench :663-679 (``P`` an operator: ``P(y)``; a plain array is the reference's diagonal)."""
    if not callable(P):
        P = _Diagonal(P)
    y = g - g_prev
    yp = np.vdot(y, p_prev)
    if not abs(yp) > 1.0e-12:
        return np.inf
    Py = P(y)
    return np.vdot(g, Py) / yp - (np.vdot(y, Py) / yp) * (np.vdot(p_prev, g) / yp)


def initial_alpha(gp, pHp, overstep):
    """bench :606-610."""
    with np.errstate(divide="ignore", invalid="ignore"):
        alpha = np.float64(-gp) / np.float64(pHp)
    if np.isnan(alpha) or alpha == -np.inf:
        alpha = 0.0
    elif alpha == np.inf:
        alpha = 1.0
    if not (alpha > 0.0 and np.isfinite(alpha)):
        alpha = 1.0
    return alpha * overstep


def minimize(
    problem,
    x0,
    *,
    max_steps=1500,
    overstep=1.0,
    max_backtracking_steps=8,
    armijo=1.0e-4,
    rtol_grad=0.0,
    history=None,
    block_jacobi=False,
    psd=False,
):
    """PNCG loop, bench :254-329 with the fused operator calls spelled out per appendix B.

    Returns ``(x, info)``.  ``rtol_grad`` terminates on ``|g| <= rtol_grad * |g_0|``."""
    x = np.array(x0, copy=True)
    g_prev = np.zeros_like(x)
    p_prev = np.zeros_like(x)
    f = problem.fun(x)
    g = problem.grad(x)
    g0_norm = np.linalg.norm(g)
    n_accepted = 0
    k = 0
    for k in range(max_steps):
        if k > 0:
            f = problem.fun(x)
            g = problem.grad(x)
        gnorm = np.linalg.norm(g)
        if history is not None:
            history.append((k, float(f), float(gnorm)))
        if gnorm <= rtol_grad * g0_norm:
            break
        P = BlockJacobi(problem, x, psd) if block_jacobi else _Diagonal(make_preconditioner(problem.hess_diag(x)))
        beta = dai_kou_beta(g, g_prev, p_prev, P)
        if k == 0 or not np.isfinite(beta) or abs(beta) > 10.0:  # bench :288-289
            beta = 0.0
        steepest = -P(g)
        p = steepest + beta * p_prev
        gp = np.vdot(g, p)
        if not (np.isfinite(gp) and gp < 0.0):  # bench :292-303
            beta, p = 0.0, steepest
            gp = np.vdot(g, p)
        pHp = problem.hess_quad_psd(x, p) if psd else problem.hess_quad(x, p)
        alpha = initial_alpha(gp, pHp, overstep)
        alpha = min(alpha, problem.max_step_size(x, p))  # _problem.py:29-34
        # bench :413-456
        tries = 0
        while True:
            x_trial = x + alpha * p
            f_trial = problem.fun(x_trial)
            accepted = np.isfinite(f_trial) and f_trial <= f + armijo * alpha * gp
            if accepted or tries >= max_backtracking_steps or not alpha > 0.0:
                break
            alpha *= 0.5
            tries += 1
        if accepted:  # bench :320-322
            x = x_trial
            n_accepted += 1
        g_prev, p_prev = g, p
    return x, {"n_steps": k + 1, "n_accepted": n_accepted, "fun": float(problem.fun(x))}
