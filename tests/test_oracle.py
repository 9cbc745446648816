"""CPU: pins the oracle -- finite differences of its own energy, the reference's known-answer
test, the derived golden energies of SURVEY.md section 8(c), and structural identities."""

import numpy as np
import pytest

from helpers import KINDS, make_case, oracle_potential
from oracle import fem, pncg, region


@pytest.fixture(scope="module")
def small_case():
    return make_case(n=2, seed=7, amp=0.2)


def _model(kind, mesh, **kw):
    pot = oracle_potential(kind, mesh)
    for k, v in kw.items():
        setattr(pot, k, v)
    return fem.Model([pot], mesh.n_points), pot


@pytest.mark.parametrize("kind", KINDS)
def test_gradient_is_derivative_of_energy(small_case, kind):
    mesh, u, p = small_case
    m, _ = _model(kind, mesh)
    g = m.grad(u)
    h = 1e-6
    rng = np.random.default_rng(0)
    for _ in range(5):
        d = rng.standard_normal(u.shape)
        fd = (m.fun(u + h * d) - m.fun(u - h * d)) / (2 * h)
        assert abs(fd - (g * d).sum()) <= 1e-7 * max(1.0, abs(fd))


@pytest.mark.parametrize("kind", KINDS)
def test_hess_prod_is_derivative_of_gradient(small_case, kind):
    mesh, u, p = small_case
    m, pot = _model(kind, mesh)
    if kind == "arap":
        pot.clamp_lambda = False  # the clamped twist eigenvalues are a PSD surrogate, not the Hessian
    Hp = m.hess_prod(u, p)
    h = 1e-6
    fd = (m.grad(u + h * p) - m.grad(u - h * p)) / (2 * h)
    assert np.abs(Hp - fd).max() <= 1e-6 * np.abs(Hp).max()


@pytest.mark.parametrize("kind", KINDS)
def test_quad_and_diag_are_consistent_with_hess_prod(small_case, kind):
    mesh, u, p = small_case
    m, pot = _model(kind, mesh, clamp_hess_diag=False, clamp_hess_quad=False)
    Hp = m.hess_prod(u, p)
    assert abs(m.hess_quad(u, p) - (p * Hp).sum()) <= 1e-10 * abs((p * Hp).sum())
    d = m.hess_diag(u)
    V = mesh.n_points
    rng = np.random.default_rng(1)
    for i in rng.choice(V, 6, replace=False):
        for c in range(3):
            e = np.zeros_like(u); e[i, c] = 1.0
            assert abs(m.hess_prod(u, e)[i, c] - d[i, c]) <= 1e-10 * np.abs(d).max()


def test_clamps_match_reference_semantics(small_case):
    """hess_diag is clamped per entry (warp/fem/_base.py:317-320), hess_quad per cell (:379-380)."""
    mesh, u, p = small_case
    pot = oracle_potential("snh", mesh)
    big = 3.0 * u  # strongly deformed: some cells have indefinite Hessians
    q = pot.elem_hess_quad(big, p)
    assert (q >= 0).all()
    pot.clamp_hess_quad = False
    q_raw = pot.elem_hess_quad(big, p)
    np.testing.assert_allclose(q, np.maximum(q_raw, 0.0))
    assert (pot.elem_hess_diag(big) >= 0).all()


def test_arap_clamped_hessian_is_psd_and_literal_bug_differs(small_case):
    mesh, u, p = small_case
    pot = oracle_potential("arap", mesh)
    m = fem.Model([pot], mesh.n_points)
    assert (p * m.hess_prod(3 * u, p)).sum() >= 0
    bug = fem.Arap(pot.cells, pot.dhdX, pot.dV, mu=pot.materials["mu"], literal_reference_bug=True)
    out = np.zeros_like(u); bug.hess_prod(u, p, out)
    assert np.abs(out - m.hess_prod(u, p)).max() > 1e-3 * np.abs(out).max()  # the swap is not the HVP


@pytest.mark.parametrize("kind", ["snh", "arap"])
def test_rigid_motion_invariance(small_case, kind):
    mesh, _, _ = small_case
    m, _ = _model(kind, mesh)
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    u = mesh.points @ R.T - mesh.points + np.array([0.3, -0.2, 0.1])
    assert abs(m.fun(u)) < 1e-10 * max(1.0, np.abs(mesh.cell_data["mu"]).max())
    assert np.abs(m.grad(u)).max() < 1e-8 * np.abs(mesh.cell_data["mu"]).max()


def test_region_layout():
    """dhdX rows 1..3 are Dm^-1, row 0 is minus their sum; dV = det/6 (SURVEY.md section 8 header)."""
    mesh, _, _ = make_case(n=2, seed=0)
    dhdX, dV = region.compute_grad(mesh.points, mesh.cells)
    X = mesh.points[mesh.cells]
    Dm = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]], axis=2)
    np.testing.assert_allclose(dhdX[:, 1:], np.linalg.inv(Dm), atol=1e-10)
    np.testing.assert_allclose(dhdX.sum(axis=1), 0.0, atol=1e-10)
    np.testing.assert_allclose(dV, np.linalg.det(Dm) / 6.0)
    assert (dV > 0).all() and abs(dV.sum() - 1.0) < 1e-12


def _kat():
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0.25, 0.25, 0.25]], float)
    cells = np.array([[0, 1, 2, 4], [0, 1, 4, 3], [0, 4, 2, 3], [4, 1, 2, 3]])
    dhdX, dV = region.compute_grad(pts, cells)
    mask = np.zeros((5, 3), bool); mask[:4] = True
    val = np.zeros((5, 3)); val[3] = [0.2, -0.1, 0.15]
    model = fem.Model([fem.Arap(cells, dhdX, dV, mu=np.ones(4))], 5)
    return pncg.ForwardProblem(model, region.DofMap(mask, val)), dV


def test_known_answer_test_of_the_reference():
    """tests/forward/test_static_simulation.py:16-93: u[4] == (0.05, -0.025, 0.0375) atol 1e-8."""
    problem, dV = _kat()
    np.testing.assert_allclose(dV, 1.0 / 24.0)
    assert problem.dof_map.n_free == 3
    e0 = problem.fun(np.zeros(3))
    x, info = pncg.minimize(problem, np.zeros(3), max_steps=1500, rtol_grad=1e-7)
    assert info["fun"] < e0
    np.testing.assert_allclose(x, [0.05, -0.025, 0.0375], atol=1e-8)
    # derived goldens, SURVEY.md section 8(c)
    assert abs(e0 - 0.010475610401894953) < 1e-15
    assert abs(info["fun"] - 0.004108894646378555) < 1e-13
    u_full = problem.dof_map.to_full(x)
    np.testing.assert_allclose(u_full[3], [0.2, -0.1, 0.15])


def test_pncg_energy_is_monotone():
    problem, _ = _kat()
    hist = []
    pncg.minimize(problem, np.zeros(3), max_steps=12, history=hist)
    f = [h[1] for h in hist]
    assert all(b <= a + 1e-16 for a, b in zip(f, f[1:]))


def test_dof_map_round_trip():
    rng = np.random.default_rng(0)
    mask = rng.random((11, 3)) < 0.4
    val = rng.standard_normal((11, 3))
    dm = region.DofMap(mask, val)
    free = rng.standard_normal(dm.n_free)
    full = dm.to_full(free)
    np.testing.assert_array_equal(full[mask], val[mask])
    np.testing.assert_array_equal(dm.to_free(full), free)
    assert (dm.to_full_grad(free)[mask] == 0).all()
