"""Jacobi-preconditioned conjugate gradients on ``hess_prod`` (the adjoint / sensitivity solve).

SURVEY.md section 3.4 and 8f rank 1: the reference's inverse problems solve ``H p = -dL/du`` on the free
DOFs with ``jax.scipy.sparse.linalg.cg(A, b, tol=1e-5, atol=1e-15, maxiter=n_free // 10, M=1/hess_diag)``
where ``A`` is ``model.hess_prod`` at a fixed ``u``
(``exp/2025/09/24/inverse-grin/src/35-inverse-small-reg.py:223-260``).  This is that solver over the same
``Problem`` protocol: the matvec is the CUDA ``hess_prod`` kernel (one pass over the elements per
iteration), the handful of vector operations between matvecs are torch ops on the problem's device with
the scalars kept on the device (one host read per ``check_every`` iterations).
"""

from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass
class PcgInfo:
    n_iters: int
    residual_norm: float
    rhs_norm: float
    converged: bool


def pcg(matvec, b: torch.Tensor, M_inv: torch.Tensor | None = None, *, x0: torch.Tensor | None = None,
        tol: float = 1.0e-5, atol: float = 1.0e-15, maxiter: int | None = None, check_every: int = 8):
    """Solves ``A x = b`` for symmetric positive definite ``A`` given as ``matvec``; ``M_inv`` is the
    (diagonal) preconditioner applied as an element-wise product.  Stops when
    ``|r| <= max(tol * |b|, atol)`` (jax.scipy.sparse.linalg.cg semantics)."""
    n = b.numel()
    if maxiter is None:
        maxiter = max(n // 10, 10)
    x = torch.zeros_like(b) if x0 is None else x0.clone()
    r = b - matvec(x) if x0 is not None else b.clone()
    z = r * M_inv if M_inv is not None else r
    p = z.clone()
    rz = torch.dot(r, z)
    b_norm = float(torch.linalg.vector_norm(b))
    target = max(tol * b_norm, atol)
    it = 0
    res = float(torch.linalg.vector_norm(r))
    while it < maxiter and res > target:
        for _ in range(min(check_every, maxiter - it)):
            Ap = matvec(p)
            alpha = rz / torch.dot(p, Ap)
            x = x + alpha * p
            r = r - alpha * Ap
            z = r * M_inv if M_inv is not None else r
            rz_new = torch.dot(r, z)
            p = z + (rz_new / rz) * p
            rz = rz_new
            it += 1
        res = float(torch.linalg.vector_norm(r))   # the only host read
        if not res == res:
            break
    return x, PcgInfo(n_iters=it, residual_norm=res, rhs_norm=b_norm, converged=res <= target)


def adjoint_solve(problem, state, rhs: torch.Tensor, **kw):
    """``H(u) p = rhs`` on the free DOFs of ``problem`` at ``state`` with the Jacobi preconditioner built from
    ``problem.hess_diag`` (non-positive entries replaced by the mean of the positive ones, like the PNCG
    preconditioner)."""
    d = problem.hess_diag(state).abs()
    pos = d > 0
    mean = d[pos].mean() if bool(pos.any()) else torch.ones((), dtype=d.dtype, device=d.device)
    M_inv = 1.0 / torch.where(pos, d, mean)
    return pcg(lambda v: problem.hess_prod(state, v), rhs, M_inv, **kw)
