"""CPU checks of the opt-in supersets (3x3 block Jacobi, analytic PSD projection): the product's per-tet math header
compiled for the host against the brute-force oracle of oracle/hessian.py, and the oracle's own invariants."""

import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from helpers import KINDS, make_case, oracle_potential
from oracle import hessian as ohess

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def host_math(tmp_path_factory):
    out = tmp_path_factory.mktemp("native") / "elem_host.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(out), str(ROOT / "tests" / "native" / "elem_host.cpp")],
                   check=True)
    return ctypes.CDLL(str(out))


def _record(ora, kind, dtype):
    T = ora.cells.shape[0]
    rec = np.zeros((T, 18 if kind == "muscle" else 12), dtype)
    rec[:, :9] = ora.dhdX[:, 1:4].reshape(T, 9)
    rec[:, 9] = ora.dV
    rec[:, 10] = ora.materials["mu"]
    if kind != "arap":
        rec[:, 11] = ora.materials["lambda_"]
    if kind == "muscle":
        rec[:, 12:] = ora.materials["activation"]
    return rec


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_dense_hessians_are_consistent(kind):
    """elem_hessian is symmetric, its block diagonals reproduce hess_diag (the clamp of _base.py:317-320 is inactive:
    every entry is analytically >= 0) and H p reproduces hess_prod; the projected Hessians are PSD and differ from
    the true ones exactly on the indefinite elements."""
    mesh, u, p = make_case(n=4, seed=2, amp=0.6)
    ora = oracle_potential(kind, mesh)
    V, T = mesh.n_points, mesh.n_cells
    H = ohess.elem_hessian(ora, u)
    assert np.abs(H - np.swapaxes(H, 1, 2)).max() < 1e-9 * np.abs(H).max()
    blocks = ohess.vertex_blocks(ora, u, V)
    d = np.zeros((V, 3)); ora.hess_diag(u, d)
    assert np.abs(np.stack([blocks[:, 0, 0], blocks[:, 1, 1], blocks[:, 2, 2]], 1) - d).max() < 1e-10 * np.abs(d).max()
    hp = np.zeros((V, 3)); ora.hess_prod(u, p, hp)
    assert np.abs(ohess.hess_prod(ora, u, p, V) - hp).max() < 1e-10 * np.abs(hp).max()
    Hp = ohess.elem_hessian_psd(ora, u)
    w, wp = np.linalg.eigvalsh(H), np.linalg.eigvalsh(Hp)
    scale = np.abs(w).max(1)
    assert (wp[:, 0] > -1e-10 * scale).all()
    indefinite = w[:, 0] < -1e-9 * scale
    changed = np.abs(H - Hp).reshape(T, -1).max(1) > 1e-9 * scale
    if kind == "arap":
        assert not changed.any()                 # the clamped twist rates already are the projection
    else:
        assert indefinite.sum() > T // 4 and (changed == indefinite).all()


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-10), (np.float32, 3e-5)], ids=["f64", "f32"])
@pytest.mark.parametrize("amp", [0.15, 1.5])
@pytest.mark.parametrize("kind", KINDS)
def test_block_and_psd_math_header_matches_dense_oracle(host_math, kind, dtype, tol, amp):
    """csrc/elem_math.cuh: vertex blocks (HESS_DIAG | HESS_OFFD), and with APL_OP_PSD the ANALYTIC eigen-projection
    (twist / flip / scaling modes in the SVD frame) of blocks, products and quadratic forms, element by element against
    the NUMERICAL projection (numpy eigh of the 9x9 d2Psi/dF2) of oracle/hessian.py."""
    mesh, u, p = make_case(n=4, seed=2, amp=amp)
    ora = oracle_potential(kind, mesh)
    T = mesh.n_cells
    rec = _record(ora, kind, dtype)
    uc = np.ascontiguousarray(u[mesh.cells], dtype); pc = np.ascontiguousarray(p[mesh.cells], dtype)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    H = ohess.elem_hessian(ora, u)
    Hp = ohess.elem_hessian_psd(ora, u)
    for sel, HH in ((0, H), (1, Hp)):
        quad = np.zeros(T, dtype); dg = np.zeros((T, 4, 3), dtype); od = np.zeros((T, 4, 3), dtype)
        host_math.elem_eval_ops_host(KINDS.index(kind), int(dtype == np.float64), sel, T, P(rec), P(uc), P(pc), P(quad), P(dg), P(od))
        HH = HH.reshape(T, 4, 3, 4, 3)
        blk = np.stack([HH[:, a, :, a, :] for a in range(4)], 1)
        rdiag = np.maximum(np.stack([blk[..., 0, 0], blk[..., 1, 1], blk[..., 2, 2]], -1), 0.0)
        roff = np.stack([blk[..., 0, 1], blk[..., 0, 2], blk[..., 1, 2]], -1)
        sc = np.abs(blk).reshape(T, -1).max(1)[:, None, None]
        assert (np.abs(dg - rdiag) / sc).max() < tol, sel
        assert (np.abs(od - roff) / sc).max() < tol, sel
    quad = np.zeros(T, dtype); dg = np.zeros((T, 4, 3), dtype); hp = np.zeros((T, 4, 3), dtype)
    host_math.elem_eval_ops_host(KINDS.index(kind), int(dtype == np.float64), 2, T, P(rec), P(uc), P(pc), P(quad), P(dg), P(hp))
    pcv = p[mesh.cells].reshape(T, 12)
    rhp = np.einsum("tij,tj->ti", Hp, pcv).reshape(T, 4, 3)
    rq = np.maximum(np.einsum("ti,tij,tj->t", pcv, Hp, pcv), 0.0)
    assert (np.abs(hp - rhp).reshape(T, -1).max(1) / np.abs(rhp).reshape(T, -1).max(1)).max() < tol
    assert (np.abs(quad - rq) / np.abs(rq).max()).max() < tol


def test_oracle_block_jacobi_pncg_reaches_the_same_minimiser():
    """The oracle's PNCG with the 3x3 block preconditioner (and with the PSD-projected passes) converges to the same
    minimiser as the reference-style scalar Jacobi.  On these cube meshes the vertex blocks are nearly isotropic, so the
    iteration counts are about equal (191 vs 194 to rtol 1e-5 at n = 4): the opt-in is a robustness feature, not a
    speed-up, here -- which is why the test does not assert fewer iterations."""
    from oracle import fem as ofem, pncg as opncg, region as oregion

    mesh, _, _ = make_case(n=4, seed=11, grading=1.0)
    mesh.cell_data.pop("Fraction")
    mesh.cell_data["lambda"] = 4.0 * mesh.cell_data["mu"]
    V = mesh.n_points
    fixed = np.zeros((V, 3), dtype=bool); fixed[mesh.points[:, 2] == 0.0] = True
    fixed[mesh.points[:, 0] == 0.0, 0] = True          # a sliding wall: vertices with mixed free / fixed components
    idx = np.flatnonzero(~fixed[:, 2])
    force = np.zeros((idx.size, 3)); force[:, 0] = 0.3; force[:, 2] = -0.2
    pots = [oracle_potential("snh", mesh), ofem.ExternalForce(force / V, idx)]
    problem = opncg.ForwardProblem(ofem.Model(pots, V), oregion.DofMap(fixed, np.zeros((V, 3))))
    x0 = np.zeros(problem.dof_map.n_free)
    runs = {}
    for name, kw in (("jacobi", {}), ("block", {"block_jacobi": True}), ("block+psd", {"block_jacobi": True, "psd": True})):
        hist = []
        x, info = opncg.minimize(problem, x0, max_steps=300, rtol_grad=1e-5, history=hist, **kw)
        assert info["n_steps"] < 300 and all(b[1] <= a[1] + 1e-15 for a, b in zip(hist, hist[1:])), name
        runs[name] = (x, info)
    for name in ("block", "block+psd"):
        assert np.abs(runs[name][0] - runs["jacobi"][0]).max() < 1e-4 * np.abs(runs["jacobi"][0]).max()
        assert abs(runs[name][1]["n_steps"] - runs["jacobi"][1]["n_steps"]) < 40
