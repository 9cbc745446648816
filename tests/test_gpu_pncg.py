"""GPU: the reference's known-answer test re-hosted, and PNCG solves against the oracle."""

import numpy as np
import pytest
import torch

from helpers import make_case, oracle_potential, rel_err

pytestmark = pytest.mark.gpu


def _kat_model(dtype, **pncg_kw):
    from apple_b200.common import FIXED_MASK, FIXED_VALUE, MU
    from apple_b200.forward import Forward, ModelBuilder
    from apple_b200.mesh import embedded_tetra_mesh
    from apple_b200.optim import PNCG
    from apple_b200.warp.fem import Arap

    mesh = embedded_tetra_mesh()
    mesh.cell_data[MU.vtk] = np.ones(mesh.n_cells)
    builder = ModelBuilder()
    builder.add_vertices(mesh)
    fixed_mask = np.zeros((mesh.n_points, 3), dtype=bool)
    fixed_value = np.zeros((mesh.n_points, 3))
    fixed_mask[:4, :] = True
    fixed_value[3] = np.array([0.2, -0.1, 0.15])
    mesh.point_data[FIXED_MASK.vtk] = fixed_mask
    mesh.point_data[FIXED_VALUE.vtk] = fixed_value
    builder.add_fixed(mesh)
    builder.add_potential(Arap.from_pyvista(mesh, dtype=dtype))
    model = builder.finalize()
    optimizer = PNCG(**pncg_kw) if pncg_kw else None
    return model, Forward(model, optimizer=optimizer), fixed_value


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "generic"])
def test_forward_static_simulation_end_to_end(native_lib, fused):
    """tests/forward/test_static_simulation.py:58-93 of the reference, same asserts (fp64)."""
    model, forward, fixed_value = _kat_model(torch.float64, fused=fused)
    initial_energy = float(forward.problem.fun(forward.state))
    solution = forward.step()
    final_energy = float(forward.problem.fun(forward.state))

    assert solution.result.name == "PRIMARY_SUCCESS"
    assert solution.stats["fused"] == fused
    assert model.n_free == 3
    assert np.isfinite(initial_energy)
    assert np.isfinite(final_energy)
    assert final_energy < initial_energy
    # derived goldens
    assert abs(initial_energy - 0.010475610401894953) < 1e-14
    assert abs(final_energy - 0.004108894646378555) < 1e-12

    u_full = forward.state.u.cpu().numpy()
    np.testing.assert_allclose(u_full[:4], fixed_value[:4])
    np.testing.assert_allclose(u_full[4], [0.05, -0.025, 0.0375], atol=1e-8)


def test_stepping_protocol_and_state_fields(native_lib):
    """init / step / terminate / postprocess with the state fields the reference's drivers log."""
    from apple_b200.optim import Result

    model, forward, _ = _kat_model(torch.float64)
    problem, state = forward.problem, forward.state
    opt_state = forward.optimizer.init(problem, state, forward.free)
    result = Result.UNKNOWN_ERROR
    energies = []
    for _ in range(100):
        state, opt_state = forward.optimizer.step(problem, state, opt_state)
        ls, cs, hd = opt_state.line_search_state, opt_state.convergence_state, opt_state.hess_damping_state
        assert np.isfinite(ls.alpha) and np.isfinite(ls.f_alpha) and isinstance(ls.ok, bool)
        assert cs.grad_norm_first > 0 and hd.hess_diag_mean > 0
        assert opt_state.direction.shape == (3,)
        energies.append(ls.f_alpha)
        done, result = forward.optimizer.terminate(problem, state, opt_state)
        if done:
            break
    solution = forward.optimizer.postprocess(problem, state, opt_state, result)
    assert solution.success and result is Result.PRIMARY_SUCCESS
    assert all(b <= a + 1e-15 for a, b in zip(energies, energies[1:]))  # monotone energy


def _cube_problem(dtype, n=8, kinds=("snh",), gravity=True, spd=False):
    from apple_b200.common import FIXED_MASK, FIXED_VALUE
    from apple_b200.forward import ModelBuilder
    from apple_b200.mesh import lumped_vertex_volume
    from apple_b200.warp.fem import Arap, StableNeoHookean
    from apple_b200.warp.potential import ExternalForce
    from oracle import fem as ofem, pncg as opncg, region as oregion

    mesh, _, _ = make_case(n=n, seed=11, grading=1.0)
    mesh.cell_data.pop("Fraction")
    if spd:  # lambda > mu / 3 everywhere: the SNH Hessian is positive definite at the rest state
        mesh.cell_data["lambda"] = 4.0 * mesh.cell_data["mu"]
    V = mesh.n_points
    builder = ModelBuilder()
    builder.add_vertices(mesh)
    fixed = np.zeros((V, 3), dtype=bool); fixed[mesh.points[:, 2] == 0.0] = True
    mesh.point_data[FIXED_MASK.vtk] = fixed
    mesh.point_data[FIXED_VALUE.vtk] = np.zeros((V, 3))
    builder.add_fixed(mesh)
    cls = {"snh": StableNeoHookean, "arap": Arap}
    opots = []
    for k in kinds:
        builder.add_potential(cls[k].from_pyvista(mesh, dtype=dtype, name=k))
        opots.append(oracle_potential(k, mesh))
    if gravity:
        idx = np.flatnonzero(~fixed[:, 0])
        force = np.zeros((idx.size, 3)); force[:, 0] = 40.0 * lumped_vertex_volume(mesh)[idx]
        builder.add_potential(ExternalForce(idx, force, dtype=dtype, name="gravity"))
        opots.append(ofem.ExternalForce(force, idx))
    model = builder.finalize()
    oproblem = opncg.ForwardProblem(ofem.Model(opots, V), oregion.DofMap(fixed, np.zeros((V, 3))))
    return model, oproblem


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-8), (torch.float32, 1e-4)], ids=["f64", "f32"])
def test_fixed_iteration_count_matches_oracle(native_lib, dtype, tol):
    """Config-1 style solve (SNH cube, fixed base, body force): displacements after a fixed number
    of PNCG iterations agree with the oracle's PNCG within 1e-4 relative (north star)."""
    from apple_b200.forward import Forward
    from apple_b200.optim import PNCG
    from apple_b200.optim.pncg import ConvergenceCriteria
    from oracle import pncg as opncg

    model, oproblem = _cube_problem(dtype)
    iters = 40
    crit = ConvergenceCriteria(max_steps=iters, target_relative_gradient_norm=0.0)
    forward = Forward(model, optimizer=PNCG(criteria=crit, check_every=8))
    solution = forward.step()
    assert solution.stats["n_steps"] == iters and solution.stats["fused"]
    x_ref, info = opncg.minimize(oproblem, np.zeros(oproblem.dof_map.n_free), max_steps=iters)
    u_ref = oproblem.dof_map.to_full(x_ref)
    assert rel_err(forward.state.u.cpu(), u_ref) < tol
    assert rel_err(solution.stats["fun"], info["fun"]) < 10 * tol
    assert solution.stats["n_accepted"] == info["n_accepted"]


def test_generic_and_fused_paths_agree(native_lib):
    from apple_b200.forward import Forward
    from apple_b200.optim import PNCG
    from apple_b200.optim.pncg import ConvergenceCriteria

    model, _ = _cube_problem(torch.float64, n=5, kinds=("snh", "arap"))
    crit = ConvergenceCriteria(max_steps=25, target_relative_gradient_norm=0.0)
    a = Forward(model, optimizer=PNCG(criteria=crit, fused=True)); a.step()
    b = Forward(model, optimizer=PNCG(criteria=crit, fused=False)); b.step()
    assert rel_err(a.state.u.cpu(), b.state.u.cpu()) < 1e-9


def test_graph_replay_equals_eager_launches(native_lib):
    """Plain launches, the static CUDA graph and the graph with a conditional WHILE node for the line
    search run the same recurrences."""
    from apple_b200.forward import Forward
    from apple_b200.optim import PNCG
    from apple_b200.optim.pncg import ConvergenceCriteria

    model, _ = _cube_problem(torch.float32, n=6)
    crit = ConvergenceCriteria(max_steps=30, target_relative_gradient_norm=0.0)
    runs = {}
    for mode, every in ((0, 7), (1, 30), (2, 30)):
        f = Forward(model, optimizer=PNCG(criteria=crit, use_graph=mode, check_every=every))
        runs[mode] = (f, f.step())
    for mode in (1, 2):
        assert runs[mode][1].stats["n_steps"] == runs[0][1].stats["n_steps"] == 30
        assert runs[mode][1].stats["n_accepted"] == runs[0][1].stats["n_accepted"]
        assert rel_err(runs[mode][0].state.u.cpu(), runs[0][0].state.u.cpu()) < 1e-4


@pytest.mark.parametrize("mode", [0, 1, 2], ids=["eager", "graph", "graph_while"])
def test_backtracking_line_search_matches_host_driven_pncg(native_lib, mode):
    """A strict sufficient-decrease constant (c1 = 0.9) makes the Newton step fail the Armijo test, so
    trials are halved: the device-side line search (flag-guarded launches or the WHILE node) must take
    the same decisions as the host-driven generic PNCG."""
    from apple_b200.forward import Forward
    from apple_b200.optim import PNCG
    from apple_b200.optim.pncg import ConvergenceCriteria, LineSearch

    model, _ = _cube_problem(torch.float64, n=5, kinds=("snh", "arap"))
    crit = ConvergenceCriteria(max_steps=12, target_relative_gradient_norm=0.0)
    ls = LineSearch(armijo=0.9)
    a = Forward(model, optimizer=PNCG(criteria=crit, line_search=ls, fused=True, use_graph=mode, check_every=1))
    b = Forward(model, optimizer=PNCG(criteria=crit, line_search=ls, fused=False))
    halvings = []
    problem, state = a.problem, a.state
    opt_state = a.optimizer.init(problem, state, a.free)
    for _ in range(12):
        state, opt_state = a.optimizer.step(problem, state, opt_state)
        halvings.append(opt_state.line_search_state.step)
    b.step()
    assert max(halvings) >= 2                        # the loop really ran
    assert rel_err(state.u.cpu(), b.state.u.cpu()) < 1e-9


def test_adjoint_solve_on_hess_prod(native_lib):
    """Jacobi-PCG whose matvec is the CUDA hess_prod kernel: H p = b to 1e-8, checked by the residual
    evaluated with the oracle's Hessian-vector product."""
    from apple_b200.forward import Forward
    from apple_b200.optim import adjoint_solve

    model, oproblem = _cube_problem(torch.float64, n=5, kinds=("snh",), gravity=False, spd=True)
    forward = Forward(model)
    n = model.n_free
    rng = np.random.default_rng(3)
    b = rng.standard_normal(n)
    rhs = torch.as_tensor(b, device=forward.state.u.device)
    x, info = adjoint_solve(forward.problem, forward.state, rhs, tol=1e-8, maxiter=4 * n)
    assert info.converged
    res = oproblem.hess_prod(np.zeros(n), x.cpu().numpy()) - b
    assert np.linalg.norm(res) <= 1e-7 * np.linalg.norm(b)


def test_config1_cube_static_solve_under_gravity(native_lib):
    """BASELINE config 1: 16^3 x 5 = 20,480-tet Stable Neo-Hookean cube (E = 1e4..1e5, nu = 0.3..0.45),
    base fixed, gravity as an ExternalForce on the lumped vertex volumes, PNCG.  Displacements after a
    fixed number of iterations against the oracle's PNCG driven by the C restatement of the operators."""
    from apple_b200.common import FIXED_MASK, FIXED_VALUE, lame_converter
    from apple_b200.forward import Forward, ModelBuilder
    from apple_b200.mesh import cube_tet_mesh, lumped_vertex_volume
    from apple_b200.optim import PNCG
    from apple_b200.optim.pncg import ConvergenceCriteria
    from apple_b200.warp.fem import StableNeoHookean
    from apple_b200.warp.potential import ExternalForce
    from oracle import cbind, fem as ofem, pncg as opncg, region as oregion

    mesh = cube_tet_mesh(16)
    assert mesh.n_cells == 20_480 and mesh.n_points == 4_913
    rng = np.random.default_rng(0)
    la, mu = lame_converter(10.0 ** rng.uniform(4, 5, mesh.n_cells), rng.uniform(0.3, 0.45, mesh.n_cells))
    mesh.cell_data["mu"], mesh.cell_data["lambda"] = mu, la
    V = mesh.n_points
    fixed = np.zeros((V, 3), bool); fixed[mesh.points[:, 2] == 0.0] = True
    idx = np.flatnonzero(~fixed[:, 0])
    force = np.zeros((idx.size, 3)); force[:, 2] = -9.8 * 1.0e3 * lumped_vertex_volume(mesh)[idx]
    b = ModelBuilder()
    b.add_vertices(mesh)
    mesh.point_data[FIXED_MASK.vtk] = fixed
    mesh.point_data[FIXED_VALUE.vtk] = np.zeros((V, 3))
    b.add_fixed(mesh)
    b.add_potential(StableNeoHookean.from_pyvista(mesh, dtype=torch.float64, name="body"))
    b.add_potential(ExternalForce(idx, force, dtype=torch.float64, name="gravity"))
    iters = 60
    crit = ConvergenceCriteria(max_steps=iters, target_relative_gradient_norm=0.0)
    fwd = Forward(b.finalize(), optimizer=PNCG(criteria=crit, check_every=20))
    sol = fwd.step()
    dhdX, dV = oregion.compute_grad(mesh.points, mesh.cells)
    omodel = ofem.Model([cbind.CPotential("snh", mesh.cells, dhdX, dV, mu, la), ofem.ExternalForce(force, idx)], V)
    oproblem = opncg.ForwardProblem(omodel, oregion.DofMap(fixed, np.zeros((V, 3))))
    x_ref, info = opncg.minimize(oproblem, np.zeros(oproblem.dof_map.n_free), max_steps=iters)
    u_ref = oproblem.dof_map.to_full(x_ref)
    assert sol.stats["n_steps"] == iters
    assert info["fun"] < 0 and sol.stats["fun"] < 0                      # the body sags: energy decreased from 0
    assert rel_err(fwd.state.u.cpu(), u_ref) < 1e-6                      # north star: 1e-4 relative
    assert rel_err(sol.stats["fun"], info["fun"]) < 1e-8
