"""Golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the oracle):
CPU -- the oracle still reproduces them; GPU -- the CUDA path reproduces them through the C ABI."""

from pathlib import Path

import numpy as np
import pytest

from helpers import KINDS, oracle_potential, rel_err

GOLD = Path(__file__).resolve().parent / "golden"


def _golden_mesh(g):
    from apple_b200.mesh import TetMesh

    mesh = TetMesh(g["points"], g["cells"])
    for k in g.files:
        if k.startswith("cell_"):
            mesh.cell_data[k[5:]] = g[k]
    return mesh


def test_oracle_reproduces_operator_goldens():
    from oracle import fem as ofem

    g = np.load(GOLD / "operators_n3.npz")
    mesh = _golden_mesh(g)
    u, p = g["u"], g["p"]
    for kind in KINDS:
        m = ofem.Model([oracle_potential(kind, mesh)], mesh.n_points)
        assert rel_err(m.fun(u), g[f"{kind}_fun"]) < 1e-13
        assert rel_err(m.grad(u), g[f"{kind}_grad"]) < 1e-12
        assert rel_err(m.hess_diag(u), g[f"{kind}_hess_diag"]) < 1e-12
        assert rel_err(m.hess_prod(u, p), g[f"{kind}_hess_prod"]) < 1e-12
        assert rel_err(m.hess_quad(u, p), g[f"{kind}_hess_quad"]) < 1e-12


def test_known_answer_golden():
    g = np.load(GOLD / "kat_arap.npz")
    assert abs(float(g["energy_initial"]) - 0.010475610401894953) < 1e-15
    assert abs(float(g["energy_final"]) - 0.004108894646378555) < 1e-12
    np.testing.assert_allclose(g["x12"], g["expected_u4"], atol=1e-8)        # the reference's own assert
    assert (np.diff(g["pncg_energy_history"]) <= 1e-16).all()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype_name,tol", [("float64", 1e-10), ("float32", 1e-5)])
def test_cuda_reproduces_operator_goldens(native_lib, dtype_name, tol):
    import torch

    from helpers import cuda_potential

    dtype = getattr(torch, dtype_name)
    g = np.load(GOLD / "operators_n3.npz")
    mesh = _golden_mesh(g)
    V, T = mesh.n_points, mesh.n_cells
    ud = torch.as_tensor(g["u"], dtype=dtype, device="cuda").contiguous()
    pd = torch.as_tensor(g["p"], dtype=dtype, device="cuda").contiguous()
    for kind in KINDS:
        pot = cuda_potential(kind, mesh, dtype)
        fun = torch.zeros(1, dtype=dtype, device="cuda"); quad = torch.zeros(1, dtype=dtype, device="cuda")
        grad, diag, prod = (torch.zeros((V, 3), dtype=dtype, device="cuda") for _ in range(3))
        pot.eval(31, ud, pd, fun=fun, quad=quad, grad=grad, diag=diag, prod=prod)
        assert rel_err(fun.cpu(), g[f"{kind}_fun"]) < tol
        assert rel_err(quad.cpu(), g[f"{kind}_hess_quad"]) < tol
        assert rel_err(grad.cpu(), g[f"{kind}_grad"]) < tol
        assert rel_err(diag.cpu(), g[f"{kind}_hess_diag"]) < tol
        assert rel_err(prod.cpu(), g[f"{kind}_hess_prod"]) < tol
        # checksum of checksums: the assembled gradient sums to the sum of the per-element gradients
        assert abs(float(grad.sum()) - g[f"{kind}_elem_grad"].sum()) <= 50 * tol * np.abs(g[f"{kind}_elem_grad"]).sum()


@pytest.mark.gpu
def test_cuda_pncg_reproduces_cube_golden(native_lib):
    import torch

    from apple_b200.common import FIXED_MASK, FIXED_VALUE
    from apple_b200.forward import Forward, ModelBuilder
    from apple_b200.mesh import TetMesh
    from apple_b200.optim import PNCG
    from apple_b200.optim.pncg import ConvergenceCriteria
    from apple_b200.warp.fem import StableNeoHookean
    from apple_b200.warp.potential import ExternalForce

    g = np.load(GOLD / "pncg_cube_n4.npz")
    mesh = TetMesh(g["points"], g["cells"], cell_data={"mu": g["mu"], "lambda": g["lam"]})
    b = ModelBuilder()
    b.add_vertices(mesh)
    mesh.point_data[FIXED_MASK.vtk] = g["fixed_mask"]
    mesh.point_data[FIXED_VALUE.vtk] = np.zeros_like(g["points"])
    b.add_fixed(mesh)
    b.add_potential(StableNeoHookean.from_pyvista(mesh, dtype=torch.float64))
    b.add_potential(ExternalForce(g["force_index"], g["force"], dtype=torch.float64))
    crit = ConvergenceCriteria(max_steps=25, target_relative_gradient_norm=0.0)
    fwd = Forward(b.finalize(), optimizer=PNCG(criteria=crit))
    sol = fwd.step()
    assert sol.stats["n_steps"] == 25 and sol.stats["n_accepted"] == int(g["n_accepted"])
    assert rel_err(fwd.state.u.cpu(), g["u25"]) < 1e-8          # north star: 1e-4 relative
    assert rel_err(sol.stats["fun"], g["fun25"]) < 1e-10
