"""GPU parity: every operator of every energy, through the C ABI, against the oracle on the same
seeded inputs.  Tolerances are the north star's: 1e-5 relative in fp32, 1e-10 in fp64."""

import numpy as np
import pytest
import torch

from helpers import KINDS, cuda_potential, make_case, oracle_potential, rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1.0e-5, torch.float64: 1.0e-10}


@pytest.fixture(scope="module")
def case():
    return make_case(n=7, seed=3)


def _dev(a, dtype, ld=3):
    t = torch.as_tensor(a, dtype=dtype, device="cuda")
    if ld == 4:
        t = torch.cat([t, torch.full((t.shape[0], 1), 7.0, dtype=dtype, device="cuda")], dim=1)  # junk padding
    return t.contiguous()


@pytest.mark.parametrize("scatter", [0, 1, 2], ids=["tile", "atomic", "tile_simple"])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("kind", KINDS)
def test_five_operators_match_oracle(native_lib, case, kind, dtype, scatter):
    mesh, u, p = case
    V = mesh.n_points
    ora = oracle_potential(kind, mesh)
    pot = cuda_potential(kind, mesh, dtype, scatter=scatter)
    ud, pd = _dev(u, dtype), _dev(p, dtype)
    tol = TOL[dtype]

    out = torch.zeros(1, dtype=dtype, device="cuda")
    pot.fun(ud, out)
    ref = np.zeros(1); ora.fun(u, ref)
    assert rel_err(out.cpu(), ref) < tol

    out = torch.zeros(1, dtype=dtype, device="cuda")
    pot.hess_quad(ud, pd, out)
    ref = np.zeros(1); ora.hess_quad(u, p, ref)
    assert rel_err(out.cpu(), ref) < tol

    for name, args, oargs in (("grad", (ud,), (u,)), ("hess_diag", (ud,), (u,)), ("hess_prod", (ud, pd), (u, p))):
        out = torch.zeros((V, 3), dtype=dtype, device="cuda")
        getattr(pot, name)(*args, out)
        ref = np.zeros((V, 3)); getattr(ora, name)(*oargs, ref)
        assert rel_err(out.cpu(), ref) < tol, name


@pytest.mark.parametrize("scatter", [0, 2], ids=["tile", "tile_simple"])
@pytest.mark.parametrize("ld", [3, 4])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("kind", KINDS)
def test_fused_pass_matches_separate_operators(native_lib, case, kind, dtype, ld, scatter):
    """One pass computing {fun, grad, hess_diag, hess_prod} + hess_quad == five separate calls,
    for both nodal layouts (vec3 rows and 16-byte padded rows); outputs accumulate."""
    from apple_b200 import _lib

    mesh, u, p = case
    V = mesh.n_points
    ora = oracle_potential(kind, mesh)
    pot = cuda_potential(kind, mesh, dtype, scatter=scatter)
    ud, pd = _dev(u, dtype, ld), _dev(p, dtype, ld)
    fun = torch.full((1,), 2.0, dtype=dtype, device="cuda")       # accumulate semantics: start non-zero
    quad = torch.zeros(1, dtype=dtype, device="cuda")
    grad, diag, prod = (torch.zeros((V, ld), dtype=dtype, device="cuda") for _ in range(3))
    pot.eval(31, ud, pd, fun=fun, quad=quad, grad=grad, diag=diag, prod=prod)
    tol = TOL[dtype]
    e = np.zeros(1); ora.fun(u, e)
    q = np.zeros(1); ora.hess_quad(u, p, q)
    assert rel_err(fun.cpu() - 2.0, e) < 4 * tol
    assert rel_err(quad.cpu(), q) < tol
    for name, got, oargs in (("grad", grad, (u,)), ("hess_diag", diag, (u,)), ("hess_prod", prod, (u, p))):
        ref = np.zeros((V, 3)); getattr(ora, name)(*oargs, ref)
        assert rel_err(got[:, :3].cpu(), ref) < tol, name
        if ld == 4:
            assert float(got[:, 3].abs().max()) == 0.0
    # the metric kernel: fun + grad + hess_prod in one pass
    fun.zero_(); grad.zero_(); prod.zero_()
    pot.eval(_lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_PROD, ud, pd, fun=fun, grad=grad, prod=prod)
    assert rel_err(fun.cpu(), e) < tol
    ref = np.zeros((V, 3)); ora.hess_prod(u, p, ref)
    assert rel_err(prod[:, :3].cpu(), ref) < tol


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_per_element_parity(native_lib, dtype):
    """Per-element energy / gradient / HVP: disconnected tets, so the assembled fields ARE the
    per-element contributions."""
    rng = np.random.default_rng(5)
    from apple_b200.mesh import TetMesh

    n = 3000
    ref_tet = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    X = ref_tet[None] + 0.08 * rng.standard_normal((n, 4, 3))
    mesh = TetMesh(X.reshape(-1, 3), np.arange(4 * n).reshape(n, 4))
    mesh.cell_data["mu"] = rng.uniform(1, 3, n)
    mesh.cell_data["lambda"] = rng.uniform(1, 9, n)
    mesh.cell_data["activation"] = 0.1 * rng.standard_normal((n, 6))
    u = 0.05 * rng.standard_normal((4 * n, 3))
    p = rng.standard_normal((4 * n, 3))
    tol = TOL[dtype]
    for kind in KINDS:
        ora = oracle_potential(kind, mesh)
        pot = cuda_potential(kind, mesh, dtype)
        ud, pd = _dev(u, dtype), _dev(p, dtype)
        g = torch.zeros((4 * n, 3), dtype=dtype, device="cuda"); pot.grad(ud, g)
        h = torch.zeros((4 * n, 3), dtype=dtype, device="cuda"); pot.hess_prod(ud, pd, h)
        eg, eh = ora.elem_grad(u), ora.elem_hess_prod(u, p)
        g = g.cpu().numpy().reshape(n, 4, 3); h = h.cpu().numpy().reshape(n, 4, 3)
        scale_g = np.abs(eg).reshape(n, -1).max(1); scale_h = np.abs(eh).reshape(n, -1).max(1)
        assert (np.abs(g - eg).reshape(n, -1).max(1) / scale_g).max() < 20 * tol, kind
        assert (np.abs(h - eh).reshape(n, -1).max(1) / scale_h).max() < 20 * tol, kind


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_model_sum_and_external_force(native_lib, case, dtype):
    """WarpModel zeroes then sums; ExternalForce energy and gradient (config 1's gravity load)."""
    from apple_b200.mesh import lumped_vertex_volume
    from apple_b200.warp.model import WarpModel, WarpModelAdapter
    from apple_b200.warp.potential import ExternalForce
    from oracle import fem as ofem

    mesh, u, p = case
    V = mesh.n_points
    idx = np.arange(0, V, 3)
    force = np.zeros((idx.size, 3)); force[:, 2] = -9.8 * 1000.0 * lumped_vertex_volume(mesh)[idx]
    pots = {k: cuda_potential(k, mesh, dtype, name=k) for k in ("snh", "arap")}
    pots["force"] = ExternalForce(idx, force, dtype=dtype, name="force")
    adapter = WarpModelAdapter(WarpModel(pots), n_points=V)
    omodel = ofem.Model([oracle_potential("snh", mesh), oracle_potential("arap", mesh), ofem.ExternalForce(force, idx)], V)
    ud, pd = _dev(u, dtype), _dev(p, dtype)
    tol = TOL[dtype]
    assert rel_err(adapter.fun(ud).cpu(), omodel.fun(u)) < tol
    assert rel_err(adapter.grad(ud).cpu(), omodel.grad(u)) < tol
    assert rel_err(adapter.hess_diag(ud).cpu(), omodel.hess_diag(u)) < tol
    assert rel_err(adapter.hess_prod(ud, pd).cpu(), omodel.hess_prod(u, p)) < tol
    assert rel_err(adapter.hess_quad(ud, pd).cpu(), omodel.hess_quad(u, p)) < tol
    f, g, h = adapter.fun_grad_hess_prod(ud, pd)
    assert rel_err(f.cpu(), omodel.fun(u)) < tol
    assert rel_err(g.cpu(), omodel.grad(u)) < tol
    assert rel_err(h.cpu(), omodel.hess_prod(u, p)) < tol


def test_edge_cases(native_lib):
    """Single tet, ragged last tile, repeated calls (counter reset), and loud failures."""
    from apple_b200 import NativeError
    from apple_b200.mesh import TetMesh

    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    mesh = TetMesh(pts, np.array([[0, 1, 2, 3]]))
    mesh.cell_data["mu"] = np.array([2.0]); mesh.cell_data["lambda"] = np.array([5.0])
    pot = cuda_potential("snh", mesh, torch.float64)
    ora = oracle_potential("snh", mesh)
    u = 0.1 * np.random.default_rng(0).standard_normal((4, 3))
    ud = torch.as_tensor(u, device="cuda")
    for _ in range(3):  # repeated launches reuse the per-handle reduction scratch
        out = torch.zeros(1, dtype=torch.float64, device="cuda"); pot.fun(ud, out)
        ref = np.zeros(1); ora.fun(u, ref)
        assert rel_err(out.cpu(), ref) < 1e-12
    # 257 tets -> two tiles, the second with one tet
    mesh2, u2, p2 = make_case(n=4, seed=1)
    sub = TetMesh(mesh2.points, mesh2.cells[:257], cell_data={k: v[:257] for k, v in mesh2.cell_data.items()})
    pot2 = cuda_potential("arap", sub, torch.float64); ora2 = oracle_potential("arap", sub)
    assert pot2.info["n_tiles"] >= 2
    g = torch.zeros((sub.n_points, 3), dtype=torch.float64, device="cuda")
    pot2.grad(torch.as_tensor(u2, device="cuda"), g)
    ref = np.zeros((sub.n_points, 3)); ora2.grad(u2, ref)
    assert rel_err(g.cpu(), ref) < 1e-10
    # wrong dtype / CPU tensors fail loudly
    with pytest.raises((TypeError, NativeError)):
        pot.fun(ud.float(), torch.zeros(1, dtype=torch.float64, device="cuda"))
    with pytest.raises(NativeError):
        pot.fun(ud.cpu(), torch.zeros(1, dtype=torch.float64))
    # an out-of-range vertex index is rejected at setup
    bad = TetMesh(pts, np.array([[0, 1, 2, 3]])); bad.cell_data.update(mesh.cell_data)
    bad.cells[0, 3] = 9
    with pytest.raises((NativeError, IndexError)):
        cuda_potential("snh", bad, torch.float64)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("kind", KINDS)
def test_rest_state_and_isotropic_scaling(native_lib, kind, dtype):
    """Degenerate spectra: F = I (rest) and F = s I (all singular values equal) -- the states PNCG
    starts from.  Everything must be finite and match the oracle (repeated singular values make the
    individual SVD factors arbitrary but every assembled quantity is still well defined)."""
    mesh, _, p = make_case(n=5, seed=9)
    V = mesh.n_points
    ora = oracle_potential(kind, mesh)
    pot = cuda_potential(kind, mesh, dtype)
    pd = _dev(p, dtype)
    tol = 10 * TOL[dtype]
    for u in (np.zeros((V, 3)), 0.1 * mesh.points, -0.05 * mesh.points):
        ud = _dev(u, dtype)
        fun = torch.zeros(1, dtype=dtype, device="cuda"); quad = torch.zeros(1, dtype=dtype, device="cuda")
        grad, diag, prod = (torch.zeros((V, 3), dtype=dtype, device="cuda") for _ in range(3))
        pot.eval(31, ud, pd, fun=fun, quad=quad, grad=grad, diag=diag, prod=prod)
        for t in (fun, quad, grad, diag, prod):
            assert bool(torch.isfinite(t).all())
        e = np.zeros(1); ora.fun(u, e)
        q = np.zeros(1); ora.hess_quad(u, p, q)
        scale = float(np.abs(mesh.cell_data["mu"]).max())
        assert abs(float(fun) - e[0]) <= tol * max(abs(e[0]), 1e-3 * scale)
        assert rel_err(quad.cpu(), q) < tol
        for name, got, oargs in (("hess_diag", diag, (u,)), ("hess_prod", prod, (u, p))):
            ref = np.zeros((V, 3)); getattr(ora, name)(*oargs, ref)
            assert rel_err(got.cpu(), ref) < tol, name
        ref = np.zeros((V, 3)); ora.grad(u, ref)
        dref = np.zeros((V, 3)); ora.hess_diag(u, dref)
        # gradient may vanish identically (rest state): compare against the force scale diag * h
        assert np.abs(grad.cpu().numpy() - ref).max() <= tol * max(np.abs(ref).max(), np.abs(dref).max() * 0.2)


@pytest.mark.parametrize("scatter", [0, 1, 2], ids=["tile", "atomic", "tile_simple"])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_fused_snh_arap_equals_sum_of_potentials(native_lib, case, dtype, scatter):
    """Two potentials on the same cells fused into one pass == WarpModel's sum over potentials
    (warp/model/_model.py:13-36), with different Fraction / materials per potential and the clamps
    applied per potential."""
    from apple_b200.warp.fem import FusedSnhArap, fuse_potentials
    from oracle import fem as ofem

    mesh, u, p = case
    V = mesh.n_points
    m2 = mesh.copy()
    m2.cell_data["Fraction"] = 1.0 - 0.5 * mesh.cell_data["Fraction"]
    m2.cell_data["mu"] = mesh.cell_data["mu"][::-1].copy()
    snh = cuda_potential("snh", mesh, dtype, name="snh", scatter=scatter)
    arap = cuda_potential("arap", m2, dtype, name="arap", scatter=scatter)
    fused = fuse_potentials({"snh": snh, "arap": arap})
    assert list(fused) == ["snh+arap"] and isinstance(fused["snh+arap"], FusedSnhArap)
    pot = fused["snh+arap"]
    ora = ofem.Model([oracle_potential("snh", mesh), oracle_potential("arap", m2)], V)
    ud, pd = _dev(3.0 * u, dtype), _dev(p, dtype)      # large deformation: some cells have clamped p.Hp
    u3 = 3.0 * u
    fun = torch.zeros(1, dtype=dtype, device="cuda"); quad = torch.zeros(1, dtype=dtype, device="cuda")
    grad, diag, prod = (torch.zeros((V, 3), dtype=dtype, device="cuda") for _ in range(3))
    pot.eval(31, ud, pd, fun=fun, quad=quad, grad=grad, diag=diag, prod=prod)
    tol = 2 * TOL[dtype]
    assert rel_err(fun.cpu(), ora.fun(u3)) < tol
    assert rel_err(quad.cpu(), ora.hess_quad(u3, p)) < tol
    assert rel_err(grad.cpu(), ora.grad(u3)) < tol
    assert rel_err(diag.cpu(), ora.hess_diag(u3)) < tol
    assert rel_err(prod.cpu(), ora.hess_prod(u3, p)) < tol
    # potentials over different cells are left alone
    other = cuda_potential("arap", make_case(n=3, seed=1)[0], dtype, name="other")
    assert set(fuse_potentials({"snh": snh, "other": other})) == {"snh", "other"}
