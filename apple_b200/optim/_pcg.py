"""Jacobi-preconditioned conjugate gradients on ``hess_prod`` (the adjoint / sensitivity solve).

SURVEY.md section 3.4 and 8f rank 1: the reference's inverse problems solve ``H p = -dL/du`` on the free
DOFs with ``jax.scipy.sparse.linalg.cg(A, b, tol=1e-5, atol=1e-15, maxiter=n_free // 10, M=1/hess_diag)``
where ``A`` is ``model.hess_prod`` at a fixed ``u``
(``exp/2025/09/24/inverse-grin/src/35-inverse-small-reg.py:223-260``).  This is that solver over the same
``Problem`` protocol, in two forms with the same recurrences:

* FUSED (the product path, taken by ``adjoint_solve`` for a ``ForwardProblem`` over FEM potentials): the whole
  solve runs inside ``libapple_b200.so`` (``apple_b200/csrc/pcg.cu``, ``apl_pcg_*``) -- fixed DOFs masked instead of
  gathered / scattered, every scalar on the device, one iteration = five launches replayed as a CUDA graph, the host
  reads the scalars once per ``check_every`` iterations;
* GENERIC ``pcg(matvec, b, M_inv)``: any matvec callback, torch vector ops in between (compatibility path).
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


def _fused_pcg(problem, state, rhs, *, tol, atol, maxiter, check_every, psd, use_graph, x0):
    import ctypes

    from apple_b200 import _lib, config
    from apple_b200.warp.fem import WarpPotentialFem

    model = problem.model
    dm = model.dof_map
    dtype, device, n = state.u.dtype, state.u.device, model.n_points
    new = lambda: torch.zeros((n, 4), dtype=dtype, device=device)  # noqa: E731
    u, x, b, r, p, Ap, diag = (new() for _ in range(7))
    u[:, :3] = state.u
    b[:, :3] = dm.to_full_grad(rhs.to(dtype))
    if x0 is not None:
        x[:, :3] = dm.to_full_grad(x0.to(dtype))
    diag[:, :3] = model.hess_diag(state.u)
    mask = torch.zeros((n, 4), dtype=torch.uint8, device=device)
    mask[:, :3] = dm.free_mask().to(torch.uint8)
    scal = torch.zeros(_lib.PCG_NSCAL, dtype=torch.float64, device=device)
    host = torch.zeros(_lib.PCG_NSCAL, dtype=torch.float64).pin_memory()
    L = _lib.lib()
    handle = ctypes.c_void_p()
    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    _lib.check(L.apl_pcg_create(_lib.dtype_code(dtype), n, dev_index, _lib.dev_ptr(u), _lib.dev_ptr(x), _lib.dev_ptr(b),
                                _lib.dev_ptr(r), _lib.dev_ptr(p), _lib.dev_ptr(Ap), _lib.dev_ptr(diag), _lib.dev_ptr(mask),
                                _lib.dev_ptr(scal), ctypes.byref(handle)))
    try:
        for pot in model.warp_model.__wrapped__.potentials.values():
            if isinstance(pot, WarpPotentialFem):      # ExternalForce has no Hessian (warp/potential/_ext_force.py:80-90)
                if pot.dtype != dtype:
                    raise TypeError(f"potential {pot.name} is {pot.dtype}, state is {dtype}")
                _lib.check(L.apl_pcg_add_fem(handle, pot._handle))
        _lib.check(L.apl_pcg_set_params(handle, float(tol), float(atol), int(maxiter), int(bool(psd)), int(config.scatter),
                                        int(bool(use_graph))))
        with torch.cuda.device(device):
            stream = _lib.stream_ptr(device)
            _lib.check(L.apl_pcg_init(handle, int(x0 is None), stream))
            done = 0.0
            while True:
                host.copy_(scal, non_blocking=True)
                torch.cuda.current_stream(device).synchronize()
                done = float(host[5])
                if done != 0.0:
                    break
                _lib.check(L.apl_pcg_iterate(handle, int(check_every), stream))
    finally:
        L.apl_pcg_destroy(handle)
    info = PcgInfo(n_iters=int(host[6]), residual_norm=float(host[3]) ** 0.5, rhs_norm=float(host[4]) ** 0.5,
                   converged=done == 1.0)
    return dm.to_free_grad(x[:, :3].contiguous()), info


def adjoint_solve(problem, state, rhs: torch.Tensor, *, tol: float = 1.0e-5, atol: float = 1.0e-15,
                  maxiter: int | None = None, check_every: int = 8, x0: torch.Tensor | None = None, fused: bool = True,
                  psd: bool = False, use_graph: bool = True):
    """``H(u) p = rhs`` on the free DOFs of ``problem`` at ``state`` with the Jacobi preconditioner built from
    ``problem.hess_diag`` (non-positive entries replaced by the mean of the positive ones, like the PNCG
    preconditioner); defaults are the reference's (``tol=1e-5, atol=1e-15, maxiter=n_free // 10``).  ``psd=True``
    (opt-in, fused path): the matvec is the PSD-projected product."""
    from apple_b200.optim._pncg import _fused_supported

    if maxiter is None:
        maxiter = max(rhs.numel() // 10, 10)
    if fused and _fused_supported(problem):
        return _fused_pcg(problem, state, rhs, tol=tol, atol=atol, maxiter=maxiter, check_every=check_every, psd=psd,
                          use_graph=use_graph, x0=x0)
    if psd:
        raise NotImplementedError("the PSD-projected product exists on the fused path only")
    kw = dict(tol=tol, atol=atol, maxiter=maxiter, check_every=check_every, x0=x0)
    d = problem.hess_diag(state).abs()
    pos = d > 0
    mean = d[pos].mean() if bool(pos.any()) else torch.ones((), dtype=d.dtype, device=d.device)
    M_inv = 1.0 / torch.where(pos, d, mean)
    return pcg(lambda v: problem.hess_prod(state, v), rhs, M_inv, **kw)
