"""GPU parity at BASELINE.json's full config-2 size (58^3 x 5 = 975,560 tets, 205,379 vertices, SNH + ARAP):
every persistent CTA of the pipelined kernel processes many tiles here (shared-memory stages, vertex ring
and mbarrier phases are reused), which the small-mesh tests cannot exercise.  Checked against the C
restatement of the reference's kernels on the whole mesh, across the three assembly variants, and through
size-independent invariants (momentum balance, translation invariance, linearity of the HVP)."""

import numpy as np
import pytest
import torch

from helpers import cuda_potential, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def config2():
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from bench import build_mesh

    mesh, u, p = build_mesh(58)
    assert mesh.n_cells == 975_560 and mesh.n_points == 205_379
    return mesh, u, p


def _eval(pot, ops, ud, pd, dtype, V, scatter=None):
    fun = torch.zeros(1, dtype=dtype, device="cuda"); quad = torch.zeros(1, dtype=dtype, device="cuda")
    grad, diag, prod = (torch.zeros((V, 3), dtype=dtype, device="cuda") for _ in range(3))
    pot.eval(ops, ud, pd, fun=fun, quad=quad, grad=grad, diag=diag, prod=prod, scatter=scatter)
    torch.cuda.synchronize()
    return fun, quad, grad, diag, prod


@pytest.mark.parametrize("kind", ["snh", "arap"])
def test_full_size_operators_match_c_oracle(native_lib, config2, kind):
    from oracle import cbind, region as oregion

    mesh, u, p = config2
    V = mesh.n_points
    dhdX, dV = oregion.compute_grad(mesh.points, mesh.cells)
    cp = cbind.CPotential(kind, mesh.cells, dhdX, dV, mesh.cell_data["mu"], mesh.cell_data["lambda"])
    e, q = np.zeros(1), np.zeros(1)
    g, d, h = (np.zeros((V, 3)) for _ in range(3))
    cp.fun(u, e); cp.hess_quad(u, p, q); cp.grad(u, g); cp.hess_diag(u, d); cp.hess_prod(u, p, h)
    for dtype, tol in ((torch.float32, 1e-5), (torch.float64, 1e-10)):
        pot = cuda_potential(kind, mesh, dtype)
        ud = torch.as_tensor(u, dtype=dtype, device="cuda"); pd = torch.as_tensor(p, dtype=dtype, device="cuda")
        for rep in range(2):                      # twice: per-handle scratch and barriers are reused
            fun, quad, grad, diag, prod = _eval(pot, 31, ud, pd, dtype, V)
            assert rel_err(fun.cpu(), e) < tol and rel_err(quad.cpu(), q) < tol
            assert rel_err(grad.cpu(), g) < tol, (kind, dtype, rep)
            assert rel_err(diag.cpu(), d) < tol
            assert rel_err(prod.cpu(), h) < tol
        # the metric kernel and the two PNCG passes as separate launches
        fun, _, grad, _, prod = _eval(pot, 11, ud, pd, dtype, V)
        assert rel_err(grad.cpu(), g) < tol and rel_err(prod.cpu(), h) < tol and rel_err(fun.cpu(), e) < tol
        fun, _, grad, diag, _ = _eval(pot, 7, ud, pd, dtype, V)
        assert rel_err(grad.cpu(), g) < tol and rel_err(diag.cpu(), d) < tol
        _, quad, _, _, _ = _eval(pot, 16, ud, pd, dtype, V)
        assert rel_err(quad.cpu(), q) < tol


def test_full_size_fused_model_variants_and_invariants(native_lib, config2):
    from apple_b200.warp.fem import fuse_potentials

    mesh, u, p = config2
    V = mesh.n_points
    dtype = torch.float32
    pots = {k: cuda_potential(k, mesh, dtype, name=k) for k in ("snh", "arap")}
    fused = list(fuse_potentials(dict(pots)).values())[0]
    ud = torch.as_tensor(u, dtype=dtype, device="cuda"); pd = torch.as_tensor(p, dtype=dtype, device="cuda")
    # fused pass == sum of the two potentials; pipelined == simple == atomic assembly
    ref = [sum(x) for x in zip(*(_eval(pot, 31, ud, pd, dtype, V, scatter=1) for pot in pots.values()))]
    for scatter in (0, 2, 1):
        got = _eval(fused, 31, ud, pd, dtype, V, scatter=scatter)
        for a, b in zip(got, ref):
            assert rel_err(a.cpu(), b.cpu()) < 2e-5, scatter
    fun, quad, grad, diag, prod = _eval(fused, 31, ud, pd, dtype, V)
    scale_g, scale_h = float(grad.abs().sum()), float(prod.abs().sum())
    # momentum balance: internal forces and H p sum to zero over the vertices (checksum of the scatter)
    assert float(grad.double().sum(0).abs().max()) < 1e-5 * scale_g
    assert float(prod.double().sum(0).abs().max()) < 1e-5 * scale_h
    # translation invariance: a rigid shift changes neither energy nor gradient; H . (constant field) = 0
    shift = torch.tensor([0.3, -0.2, 0.1], dtype=dtype, device="cuda")
    fun2, _, grad2, _, prod2 = _eval(fused, 11, (ud + shift).contiguous(), torch.ones_like(pd) * shift, dtype, V)
    assert abs(float(fun2) - float(fun)) < 1e-3 * abs(float(fun))
    assert rel_err(grad2.cpu(), grad.cpu()) < 5e-3          # fp32 cancellation in (u_a - u_0) after the shift
    assert float(prod2.abs().max()) < 1e-4 * float(prod.abs().max())
    # linearity of the Hessian-vector product in p
    p2 = torch.roll(pd, 1, 0).contiguous()
    h1 = _eval(fused, 8, ud, pd, dtype, V)[4]; h2 = _eval(fused, 8, ud, p2, dtype, V)[4]
    h12 = _eval(fused, 8, ud, (2.0 * pd - 0.5 * p2).contiguous(), dtype, V)[4]
    assert rel_err(h12.cpu(), (2.0 * h1 - 0.5 * h2).cpu()) < 2e-5
    # diag is non-negative (clamped per entry and cell) and hess_quad >= 0
    assert float(diag.min()) >= 0.0 and float(quad) >= 0.0


def test_full_size_pncg_iterations_decrease_the_energy(native_lib, config2):
    """200 fused PNCG iterations of config 2 (fixed base): monotone energy, every Armijo test evaluated on
    the device; the WHILE-node graph and plain launches follow the same trajectory at the start."""
    from apple_b200.common import FIXED_MASK, FIXED_VALUE
    from apple_b200.forward import Forward, ModelBuilder
    from apple_b200.optim import PNCG
    from apple_b200.optim.pncg import ConvergenceCriteria

    mesh, u, _ = config2
    V = mesh.n_points
    dtype = torch.float32
    b = ModelBuilder(dtype=dtype, device="cuda")
    b.add_vertices(mesh)
    fixed = np.zeros((V, 3), bool); fixed[mesh.points[:, 2] == 0.0] = True
    mesh.point_data[FIXED_MASK.vtk] = fixed
    mesh.point_data[FIXED_VALUE.vtk] = np.zeros((V, 3))
    b.add_fixed(mesh)
    for k in ("snh", "arap"):
        b.add_potential(cuda_potential(k, mesh, dtype, name=k))
    model = b.finalize()
    u0 = np.ascontiguousarray(u); u0[fixed] = 0.0
    runs = {}
    for mode in (2, 0):
        crit = ConvergenceCriteria(max_steps=400, target_relative_gradient_norm=0.0)
        fwd = Forward(model, optimizer=PNCG(criteria=crit, use_graph=mode))
        fwd.state.u = torch.as_tensor(u0, dtype=dtype, device="cuda")
        problem, state = fwd.problem, fwd.state
        opt = fwd.optimizer.init(problem, state, fwd.free)
        energies = []
        f_prev = None
        for _ in range(10):
            state = opt.step(problem, state, 20)
            f = opt.line_search_state.f_alpha
            assert np.isfinite(f) and (f_prev is None or f <= f_prev * (1 + 1e-6))
            f_prev = f
            energies.append(f)
        assert opt.n_steps == 200 and opt.n_accepted >= 190
        assert opt.relative_grad_norm < 0.05
        runs[mode] = energies
    assert abs(runs[2][0] - runs[0][0]) <= 5e-3 * abs(runs[0][0])      # same trajectory after 20 iterations


def test_config4_muscle_model_full_size_matches_c_oracle(native_lib):
    """BASELINE.json configs[3] at its size: 74^3 x 5 = 2,026,120 tets, three potentials on one mesh in the pattern of
    /root/reference/exp/2026/05/06/toy/src/21-smas-prestrain-stable-neo-hookean-muscle.py:228-240 -- fat (Stable
    Neo-Hookean, Fraction = 1 - s), muscle (Stable Neo-Hookean x 1e3 with an activation field, Fraction = s) -- plus the
    adjoint-side operator: the model's hess_prod.  Energy / gradient / diagonal / HVP of the SUM against the C
    restatement of the reference (fp64, whole mesh)."""
    from apple_b200 import _lib
    from apple_b200.mesh import cube_tet_mesh
    from apple_b200.warp.model import WarpModel
    from oracle import cbind, region as oregion

    n = 74
    mesh = cube_tet_mesh(n, morton=True)
    T, V = mesh.n_cells, mesh.n_points
    assert T == 2_026_120
    rng = np.random.default_rng(1)
    cx = mesh.points[mesh.cells].mean(axis=1)
    s = np.clip(0.5 + 0.5 * np.sin(6.0 * cx[:, 0]) * np.cos(5.0 * cx[:, 1]), 0.05, 0.95)      # muscle fraction field
    mu = 10.0 ** rng.uniform(3.0, 4.0, T)
    la = 10.0 ** rng.uniform(3.5, 4.5, T)
    act = 0.05 * rng.standard_normal((T, 6))
    slab = np.abs(cx[:, 2] - 0.5) < 0.1                                                        # prestrained slab
    act[slab, :3] = np.array([1.2, 1.3 ** -2, 1.2]) - 1.0
    act[slab, 3:] = 0.0
    h = 1.0 / n
    X = mesh.points
    u = np.ascontiguousarray(0.05 * h * np.sin(7.0 * X[:, [1, 2, 0]] + 0.3) + 0.02 * h * rng.uniform(-1, 1, X.shape))
    p = rng.uniform(-1, 1, X.shape)
    dhdX, dV = oregion.compute_grad(mesh.points, mesh.cells)
    fat_o = cbind.CPotential("snh", mesh.cells, dhdX, (1.0 - s) * dV, mu, la)
    mus_o = cbind.CPotential("muscle", mesh.cells, dhdX, s * dV, 1e3 * mu, 1e3 * la, act)
    e = np.zeros(1)
    g, d, hp = (np.zeros((V, 3)) for _ in range(3))
    for o in (fat_o, mus_o):
        o.fun(u, e); o.grad(u, g); o.hess_diag(u, d); o.hess_prod(u, p, hp)
    for dtype, tol in ((torch.float32, 1e-5), (torch.float64, 1e-10)):
        fat_m, mus_m = mesh.copy(), mesh.copy()
        fat_m.cell_data.update({"mu": mu, "lambda": la, "Fraction": 1.0 - s})
        mus_m.cell_data.update({"mu": 1e3 * mu, "lambda": 1e3 * la, "Fraction": s, "activation": act})
        model = WarpModel({"fat": cuda_potential("snh", fat_m, dtype, name="fat"),
                           "muscle": cuda_potential("muscle", mus_m, dtype, name="muscle")})
        ud = torch.as_tensor(u, dtype=dtype, device="cuda"); pd = torch.as_tensor(p, dtype=dtype, device="cuda")
        fun = torch.zeros(1, dtype=dtype, device="cuda")
        grad, diag, prod = (torch.zeros((V, 3), dtype=dtype, device="cuda") for _ in range(3))
        model.eval(_lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_DIAG | _lib.OP_HESS_PROD, ud, pd, fun=fun, grad=grad,
                   diag=diag, prod=prod)
        torch.cuda.synchronize()
        assert rel_err(fun.cpu(), e) < tol, dtype
        assert rel_err(grad.cpu(), g) < tol, dtype
        assert rel_err(diag.cpu(), d) < tol, dtype
        assert rel_err(prod.cpu(), hp) < tol, dtype
        # the adjoint solve's matvec alone, accumulated over the potentials exactly like WarpModel.hess_prod
        out = torch.zeros((V, 3), dtype=dtype, device="cuda")
        model.hess_prod(ud, pd, out)
        torch.cuda.synchronize()
        assert rel_err(out.cpu(), hp) < tol, dtype


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 2e-5)], ids=["f64", "f32"])
def test_arap_on_compressed_and_inverted_elements(native_lib, dtype, tol):
    """ARAP away from the comfortable regime: strongly compressed, near-flat and INVERTED tets take the robust
    branch of csrc/elem_math.cuh (polar_twist falls back to the Jacobi SVD in the rotation-variant convention of
    warp/math/_rotation.py:9-13: U, V proper rotations, the smallest singular value carries the sign of det F).  The
    oracle's svd_rv restates the same convention with numpy's SVD; Warp's own wp.svd3 is unverifiable here
    (SURVEY.md appendix C.1), so this pins the library to the oracle, not to Warp, on these elements."""
    from helpers import make_case, oracle_potential

    mesh, u, p = make_case(n=6, seed=11, amp=0.05)
    V = mesh.n_points
    rng = np.random.default_rng(5)
    X = mesh.points
    # squash the cube to 15 % of its height around z = 0.5, shear it, and push one plane of vertices through its neighbours
    u = u.copy()
    u[:, 2] += -0.85 * (X[:, 2] - 0.5)
    u[:, 0] += 0.4 * X[:, 2]
    layer = np.isclose(X[:, 2], X[:, 2][np.argsort(X[:, 2])[V // 2]])
    u[layer, 2] += 0.25                                                   # inverts the tets above that plane
    u += 0.01 * rng.standard_normal(u.shape)
    ora = oracle_potential("arap", mesh)
    J = np.linalg.det(ora._F(u).reshape(-1, 3, 3))
    assert (J < 0).sum() > 20 and (np.abs(J) < 0.2).sum() > 50          # the case really contains such elements
    pot = cuda_potential("arap", mesh, dtype)
    ud = torch.as_tensor(u, dtype=dtype, device="cuda"); pd = torch.as_tensor(p, dtype=dtype, device="cuda")
    fun, quad, grad, diag, prod = _eval(pot, 31, ud, pd, dtype, V)
    e, q = np.zeros(1), np.zeros(1)
    g, d, hp = (np.zeros((V, 3)) for _ in range(3))
    ora.fun(u, e); ora.hess_quad(u, p, q); ora.grad(u, g); ora.hess_diag(u, d); ora.hess_prod(u, p, hp)
    assert rel_err(fun.cpu(), e) < tol
    assert rel_err(grad.cpu(), g) < tol
    assert rel_err(diag.cpu(), d) < 10 * tol        # the twist rates 2 / max(s_i + s_j, 2) kink where s_i + s_j = 2
    assert rel_err(prod.cpu(), hp) < 10 * tol
    assert rel_err(quad.cpu(), q) < 10 * tol


def test_config3_graded_mesh_full_size_matches_c_oracle(native_lib):
    """BASELINE.json configs[2] at its size: 117^3 x 5 = 8,008,065 tets on rectilinear coordinates graded by 1.02 per layer
    (rest volumes span three decades, dhdX differs per tet), Stable Neo-Hookean with materials over two decades.  The
    potential is built by the DEVICE setup (apl_fem_create_from_mesh: rest shape, Morton sort and planes on the GPU);
    energy / gradient / diagonal / HVP against the C restatement of the reference on the whole mesh, fp32 and fp64."""
    from apple_b200 import _lib
    from apple_b200.mesh import cube_tet_mesh
    from apple_b200.warp.fem import StableNeoHookean
    from oracle import cbind, region as oregion

    mesh = cube_tet_mesh(117, grading=1.02, morton=True)
    T, V = mesh.n_cells, mesh.n_points
    assert T == 8_008_065
    rng = np.random.default_rng(2)
    mu, la = 10.0 ** rng.uniform(3, 5, T), 10.0 ** rng.uniform(3, 5, T)
    X = mesh.points
    # a smooth field with up to 5 % strain whatever the local cell size (a displacement scaled by the finest layer would
    # strain the coarse cells by 1e-4, where the fp32 energy mu/2 (I2 - 3) - ... of ANY implementation of the reference's
    # formula is rounding noise)
    u = np.ascontiguousarray((0.05 / 7.0) * np.sin(7.0 * X[:, [1, 2, 0]] + 0.3))
    p = rng.uniform(-1, 1, X.shape)
    dhdX, dV = oregion.compute_grad(mesh.points, mesh.cells)
    assert dV.max() > 1000 * dV.min() > 0
    cp = cbind.CPotential("snh", mesh.cells, dhdX, dV, mu, la)
    e = np.zeros(1)
    g, d, hp = (np.zeros((V, 3)) for _ in range(3))
    cp.fun(u, e); cp.grad(u, g); cp.hess_diag(u, d); cp.hess_prod(u, p, hp)
    del cp, dhdX
    cells_d = torch.as_tensor(mesh.cells, device="cuda")
    points_d = torch.as_tensor(mesh.points, device="cuda")
    for dtype, tol in ((torch.float32, 1e-5), (torch.float64, 1e-10)):
        pot = StableNeoHookean.from_device_mesh(cells_d, points_d, mu=torch.as_tensor(mu, dtype=dtype, device="cuda"),
                                                lambda_=torch.as_tensor(la, dtype=dtype, device="cuda"), dtype=dtype)
        ud = torch.as_tensor(u, dtype=dtype, device="cuda"); pd = torch.as_tensor(p, dtype=dtype, device="cuda")
        fun, _, grad, diag, prod = _eval(pot, _lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_DIAG | _lib.OP_HESS_PROD, ud, pd, dtype, V)
        assert rel_err(fun.cpu(), e) < tol, dtype
        assert rel_err(grad.cpu(), g) < tol, dtype
        assert rel_err(diag.cpu(), d) < tol, dtype
        assert rel_err(prod.cpu(), hp) < tol, dtype
        fun, _, grad, _, prod = _eval(pot, 11, ud, pd, dtype, V)           # the metric kernel (packed pairs in fp32)
        assert rel_err(fun.cpu(), e) < tol and rel_err(grad.cpu(), g) < tol and rel_err(prod.cpu(), hp) < tol, dtype
        del pot
        torch.cuda.empty_cache()
