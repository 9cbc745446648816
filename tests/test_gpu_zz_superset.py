"""GPU parity of the opt-in supersets (include/apple_b200.h: APL_OP_HESS_OFFD, APL_OP_PSD,
apl_pncg_set_block_jacobi) through the C ABI against the brute-force oracle of oracle/hessian.py.  Default behaviour
(bits clear) is covered by the other suites and must be unchanged."""

import numpy as np
import pytest
import torch

from helpers import KINDS, cuda_potential, make_case, oracle_potential, rel_err
from oracle import hessian as ohess

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1.0e-5, torch.float64: 1.0e-10}


@pytest.mark.parametrize("ld", [3, 4])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("kind", KINDS)
def test_vertex_blocks_and_psd_operators_match_dense_oracle(native_lib, kind, dtype, ld):
    from apple_b200 import _lib

    mesh, u, p = make_case(n=7, seed=3, amp=0.6)
    V = mesh.n_points
    ora = oracle_potential(kind, mesh)
    pot = cuda_potential(kind, mesh, dtype)
    tol = TOL[dtype]

    def dev(a):
        t = torch.zeros((V, ld), dtype=dtype, device="cuda"); t[:, :3] = torch.as_tensor(a, dtype=dtype); return t

    ud, pd = dev(u), dev(p)
    new = lambda: torch.zeros((V, ld), dtype=dtype, device="cuda")  # noqa: E731
    for psd in (False, True):
        blocks = ohess.vertex_blocks(ora, u, V, psd=psd)
        diag, offd, grad = new(), new(), new()
        fun = torch.zeros(1, dtype=dtype, device="cuda")
        # the PNCG pass A of the block-Jacobi mode: energy + gradient + blocks in one pass
        pot.eval(_lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_DIAG | _lib.OP_HESS_OFFD | (_lib.OP_PSD if psd else 0), ud,
                 fun=fun, grad=grad, diag=diag, offd=offd)
        scale = np.abs(blocks).max()
        rd = np.stack([blocks[:, 0, 0], blocks[:, 1, 1], blocks[:, 2, 2]], 1)
        ro = np.stack([blocks[:, 0, 1], blocks[:, 0, 2], blocks[:, 1, 2]], 1)
        assert np.abs(diag[:, :3].cpu().numpy() - rd).max() < tol * scale, psd
        assert np.abs(offd[:, :3].cpu().numpy() - ro).max() < tol * scale, psd
        g = np.zeros((V, 3)); ora.grad(u, g)
        e = np.zeros(1); ora.fun(u, e)
        assert rel_err(grad[:, :3].cpu(), g) < tol and rel_err(fun.cpu(), e) < tol
        d2, o2 = new(), new()
        pot.hess_block(ud, d2, o2, psd=psd)
        assert torch.equal(d2, diag) or rel_err(d2.cpu(), diag.cpu()) < tol
        if not psd:   # the block diagonal IS hess_diag
            hd = new(); pot.hess_diag(ud, hd)
            assert rel_err(diag[:, :3].cpu(), hd[:, :3].cpu()) < tol
    # PSD diagonal alone (no p given: PNCG pass A with the scalar preconditioner and psd=True) == diagonal of the PSD blocks
    dpsd = new(); pot.eval(_lib.OP_HESS_DIAG | _lib.OP_PSD, ud, diag=dpsd)
    bp = ohess.vertex_blocks(ora, u, V, psd=True)
    assert np.abs(dpsd[:, :3].cpu().numpy() - np.stack([bp[:, 0, 0], bp[:, 1, 1], bp[:, 2, 2]], 1)).max() < tol * np.abs(bp).max()
    out = new(); pot.hess_prod_psd(ud, pd, out)
    assert rel_err(out[:, :3].cpu(), ohess.hess_prod(ora, u, p, V, psd=True)) < tol
    q = torch.zeros(1, dtype=dtype, device="cuda"); pot.hess_quad_psd(ud, pd, q)
    assert rel_err(q.cpu(), ohess.hess_quad(ora, u, p, psd=True)) < tol
    # projected products are descent-safe: p . H+ p >= 0 element-wise, so the per-cell clamp is inactive
    assert float((out[:, :3].double() * pd[:, :3].double()).sum()) == pytest.approx(float(q), rel=50 * tol)


def test_superset_bits_are_rejected_where_they_do_not_apply(native_lib):
    from apple_b200 import _lib

    mesh, u, p = make_case(n=4, seed=1)
    V = mesh.n_points
    pot = cuda_potential("snh", mesh, torch.float64)
    ud = torch.as_tensor(u, device="cuda"); pd = torch.as_tensor(p, device="cuda")
    out = torch.zeros((V, 3), dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        pot.eval(_lib.OP_HESS_OFFD | _lib.OP_HESS_PROD, ud, pd, prod=out, offd=out)
    with pytest.raises(_lib.NativeError):      # the atomic baseline has no block / PSD variant
        pot.eval(_lib.OP_HESS_DIAG | _lib.OP_HESS_OFFD, ud, diag=out, offd=out.clone(), scatter=_lib.SCATTER_ATOMIC)


def test_psd_passes_with_the_scalar_preconditioner(native_lib):
    """PNCG(psd=True) without the block preconditioner: pass A = fun | grad | PSD diagonal (no direction field), pass B =
    PSD quadratic form; same minimiser as the default run."""
    from test_gpu_pncg import _cube_problem

    from apple_b200.forward import Forward
    from apple_b200.optim import PNCG
    from apple_b200.optim.pncg import ConvergenceCriteria

    model, _ = _cube_problem(torch.float64, n=5)
    crit = ConvergenceCriteria(max_steps=400, target_relative_gradient_norm=1e-6)
    a = Forward(model, optimizer=PNCG(criteria=crit, psd=True)); sa = a.step()
    b = Forward(model, optimizer=PNCG(criteria=crit)); sb = b.step()
    assert sa.success and sb.success
    assert rel_err(a.state.u.cpu(), b.state.u.cpu()) < 1e-4


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-8), (torch.float32, 1e-4)], ids=["f64", "f32"])
@pytest.mark.parametrize("psd", [False, True], ids=["block", "block+psd"])
def test_block_jacobi_pncg_matches_oracle(native_lib, dtype, tol, psd):
    """Fixed iteration count with the 3x3 block preconditioner (and the PSD passes): displacements, energy and the
    number of accepted steps against the oracle's PNCG with oracle/pncg.py: BlockJacobi."""
    from test_gpu_pncg import _cube_problem

    from apple_b200.forward import Forward
    from apple_b200.optim import PNCG
    from apple_b200.optim.pncg import ConvergenceCriteria
    from oracle import pncg as opncg

    model, oproblem = _cube_problem(dtype, n=6)
    iters = 30
    crit = ConvergenceCriteria(max_steps=iters, target_relative_gradient_norm=0.0)
    runs = {}
    for mode in (0, 2):
        forward = Forward(model, optimizer=PNCG(criteria=crit, check_every=8, preconditioner="block", psd=psd, use_graph=mode))
        solution = forward.step()
        assert solution.stats["n_steps"] == iters and solution.stats["fused"]
        runs[mode] = (forward.state.u.cpu().numpy(), solution)
    x_ref, info = opncg.minimize(oproblem, np.zeros(oproblem.dof_map.n_free), max_steps=iters, block_jacobi=True, psd=psd)
    u_ref = oproblem.dof_map.to_full(x_ref)
    for mode in (0, 2):
        assert rel_err(runs[mode][0], u_ref) < tol, mode
        assert rel_err(runs[mode][1].stats["fun"], info["fun"]) < 10 * tol
        assert runs[mode][1].stats["n_accepted"] == info["n_accepted"]
    # and the default (scalar Jacobi) run is a different trajectory: the option really changes the preconditioner
    base = Forward(model, optimizer=PNCG(criteria=crit, check_every=8)); base.step()
    assert rel_err(base.state.u.cpu(), u_ref) > 10 * tol


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-8), (torch.float32, 1e-5)], ids=["f64", "f32"])
def test_fused_pcg_matches_host_driven_pcg_and_oracle_residual(native_lib, dtype, tol):
    """apl_pcg_* (fused adjoint solve: masked DOFs, device scalars, CUDA graph) against the generic torch-driven PCG over
    the same hess_prod kernel, and the solution's residual evaluated with the ORACLE's Hessian-vector product
    (pattern of exp/2025/09/24/inverse-grin/src/35-inverse-small-reg.py:253-260)."""
    from test_gpu_pncg import _cube_problem

    from apple_b200.forward import Forward
    from apple_b200.optim import adjoint_solve

    model, oproblem = _cube_problem(dtype, n=5, kinds=("snh",), gravity=True, spd=True)
    forward = Forward(model)
    n = model.n_free
    rng = np.random.default_rng(3)
    b = rng.standard_normal(n)
    rhs = torch.as_tensor(b, dtype=dtype, device=forward.state.u.device)
    runs = {}
    for name, kw in (("graph", dict(fused=True, use_graph=True)), ("eager", dict(fused=True, use_graph=False)),
                     ("generic", dict(fused=False))):
        x, info = adjoint_solve(forward.problem, forward.state, rhs, tol=tol, maxiter=4 * n, **kw)
        assert info.converged, name
        runs[name] = (x.double().cpu().numpy(), info)
        res = oproblem.hess_prod(np.zeros(n), runs[name][0]) - b
        assert np.linalg.norm(res) <= (10 if dtype == torch.float64 else 30) * tol * np.linalg.norm(b), name
    assert abs(runs["graph"][1].n_iters - runs["generic"][1].n_iters) <= 8 + runs["generic"][1].n_iters // 10
    # same recurrences; the REDs of hess_prod land in a different order from launch to launch, so in fp32 the iteration at
    # which |r| crosses the threshold may differ by one or two
    assert abs(runs["graph"][1].n_iters - runs["eager"][1].n_iters) <= (0 if dtype == torch.float64 else 3)
    assert rel_err(runs["graph"][0], runs["generic"][0]) < 100 * tol
    # warm start: the solution as x0 converges immediately
    x0 = torch.as_tensor(runs["graph"][0], dtype=dtype, device=rhs.device)
    _, info = adjoint_solve(forward.problem, forward.state, rhs, tol=10 * tol, maxiter=4 * n, x0=x0)
    assert info.converged and info.n_iters <= 2


def test_fused_pcg_reports_breakdown_on_indefinite_hessian_and_psd_fixes_it(native_lib):
    """At a strongly deformed state the true SNH Hessian is indefinite: CG hits a direction of non-positive curvature
    (done = 3, not converged) -- the reference falls back to NormalCG there (:262-283); with psd=True the operator is
    positive semi-definite by construction and the same solve converges."""
    from test_gpu_pncg import _cube_problem

    from apple_b200.forward import Forward
    from apple_b200.optim import adjoint_solve

    dtype = torch.float64
    model, oproblem = _cube_problem(dtype, n=5, kinds=("snh",), gravity=False)
    forward = Forward(model)
    rng = np.random.default_rng(5)
    n = model.n_free
    V = model.n_points
    u = torch.as_tensor(0.25 * rng.standard_normal((V, 3)) / 5, dtype=dtype, device="cuda")   # ~ 25 % strain noise
    import dataclasses
    state = dataclasses.replace(forward.state, u=model.dof_map.to_full(model.dof_map.to_free(u)))
    rhs = torch.as_tensor(rng.standard_normal(n), dtype=dtype, device="cuda")
    _, plain = adjoint_solve(forward.problem, state, rhs, tol=1e-6, maxiter=2 * n)
    x, proj = adjoint_solve(forward.problem, state, rhs, tol=1e-6, maxiter=20 * n, psd=True)
    assert proj.converged
    assert not plain.converged or plain.n_iters > 0     # the plain solve may break down; it must not crash or hang
    # x solves the PROJECTED system: residual with the oracle's projected product
    from oracle import hessian as ohess
    pot = oproblem.model.potentials[0]
    full = oproblem.dof_map.to_full_grad(x.cpu().numpy())
    r = oproblem.dof_map.to_free(ohess.hess_prod(pot, state.u.cpu().numpy(), full, V, psd=True)) - rhs.cpu().numpy()
    assert np.linalg.norm(r) <= 1e-5 * float(rhs.norm())


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_fused_snh_arap_supersets_equal_the_sum_of_its_parts(native_lib, dtype):
    """The fused SNH+ARAP handle (config 2 / 5's model) with the opt-in bits: vertex blocks, PSD blocks, PSD product and PSD
    quadratic form equal the sums over the two separate potentials (each validated against the oracle above)."""
    from apple_b200.warp.fem import fuse_potentials

    mesh, u, p = make_case(n=7, seed=4, amp=0.6)
    mesh.cell_data.pop("Fraction")
    V = mesh.n_points
    tol = TOL[dtype]
    parts = {k: cuda_potential(k, mesh, dtype, name=k) for k in ("snh", "arap")}
    fused = list(fuse_potentials({k: cuda_potential(k, mesh, dtype, name=k) for k in ("snh", "arap")}).values())
    assert len(fused) == 1
    fused = fused[0]
    ud = torch.as_tensor(u, dtype=dtype, device="cuda"); pd = torch.as_tensor(p, dtype=dtype, device="cuda")
    new = lambda: torch.zeros((V, 3), dtype=dtype, device="cuda")  # noqa: E731
    for psd in (False, True):
        d0, o0, d1, o1 = new(), new(), new(), new()
        fused.hess_block(ud, d0, o0, psd=psd)
        for pot in parts.values():
            pot.hess_block(ud, d1, o1, psd=psd)
        assert rel_err(d0.cpu(), d1.cpu()) < 10 * tol and rel_err(o0.cpu(), o1.cpu()) < 10 * tol, psd
    h0, h1 = new(), new()
    fused.hess_prod_psd(ud, pd, h0)
    q0, q1 = torch.zeros(1, dtype=dtype, device="cuda"), torch.zeros(1, dtype=dtype, device="cuda")
    fused.hess_quad_psd(ud, pd, q0)
    for pot in parts.values():
        pot.hess_prod_psd(ud, pd, h1)
        pot.hess_quad_psd(ud, pd, q1)
    assert rel_err(h0.cpu(), h1.cpu()) < 10 * tol
    assert rel_err(q0.cpu(), q1.cpu()) < 10 * tol
