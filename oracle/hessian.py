"""Oracle for the OPT-IN supersets of the reference's Hessian operators (test infrastructure, numpy, fp64).

The reference preconditions with the clamped scalar diagonal (``warp/fem/_base.py:307-320``) and guards the
quadratic form with a per-cell clamp (``:379-380``).  BASELINE.json's north star additionally asks for a 3x3
block-Jacobi preconditioner and an analytic per-element PSD projection.  Neither exists in the reference, so this
oracle is built from the reference's OWN device functions restated in ``oracle/fem.py`` by brute force:

* ``elem_hessian``      dense 12x12 element Hessians: the potential's ``hess_prod_func`` (``_stable_neo_hookean.py:68-83``,
                        ``_arap.py:43-58`` with the correct argument order, ...) applied to the 12 unit nodal vectors;
* ``vertex_blocks``     their four 3x3 diagonal blocks -> assembled per vertex (``diag`` of it must equal ``hess_diag``);
* ``elem_hessian_psd``  Stable Neo-Hookean kinds: the 9x9 ``d2Psi/dF2`` assembled from ``g3`` / ``_h6_matrix`` of
                        ``oracle/fem.py``, projected onto the PSD cone NUMERICALLY (``numpy.linalg.eigh``, negative
                        eigenvalues set to zero) and pulled back to the nodes, ``H_e^+ = dV B^T H_F^+ B`` with
                        ``dF = B p_cell`` (``func/_deformation.py:22-25``).  The CUDA path does the same with the
                        analytic eigen-system (``csrc/elem_math.cuh: snh_negative_modes``); agreement of the two is the test.
                        ARAP: the reference's clamped twist rates (``func/_misc.py:31-43``) already ARE the projection.
"""

from __future__ import annotations

import numpy as np

from . import fem as ofem


def _kinematics(pot, u):
    """(F, dhdX) the potential's Hessian functions are evaluated on (muscle: G = F A, dhdX A)."""
    F = pot._F(u)
    dhdX = pot.dhdX
    if isinstance(pot, ofem.StableNeoHookeanMuscle):
        A = pot._A()
        return F @ A, dhdX @ A
    return F, dhdX


def elem_hessian(pot, u) -> np.ndarray:
    """(T, 12, 12): dV * d2Psi/dx2 per element, rows / columns ordered (corner, component); no clamps."""
    T = pot.cells.shape[0]
    F = pot._F(u)
    H = np.zeros((T, 12, 12))
    for k in range(12):
        e = np.zeros((T, 4, 3))
        e[:, k // 3, k % 3] = 1.0
        H[:, :, k] = (pot.dV[:, None, None] * pot.hess_prod_func(F, e, pot.dhdX)).reshape(T, 12)
    return H


def _B(dhdX) -> np.ndarray:
    """(T, 9, 12): vec(dF) = B vec(p_cell), dF[i, J] = sum_a p[a, i] dhdX[a, J]."""
    T = dhdX.shape[0]
    B = np.zeros((T, 3, 3, 4, 3))
    for i in range(3):
        B[:, i, :, :, i] = np.swapaxes(dhdX, 1, 2)      # [J, a]
    return B.reshape(T, 9, 12)


def snh_hessian_F(pot, u) -> np.ndarray:
    """(T, 9, 9) d2Psi/dF2 of the Stable Neo-Hookean kinds at F (muscle: at G)."""
    F, _ = _kinematics(pot, u)
    T = F.shape[0]
    mu, la = pot.materials["mu"], pot.materials["lambda_"]
    g = ofem.g3(F)
    c3 = -mu + la * (ofem.I3(F) - 1.0)
    H = np.zeros((T, 9, 9))
    for k in range(9):
        E = np.zeros((T, 3, 3))
        E[:, k // 3, k % 3] = 1.0
        dP = (mu[:, None, None] * E + (la * ofem.ddot(g, E))[:, None, None] * g
              + c3[:, None, None] * ofem._h6_matrix(F, E))
        H[:, :, k] = dP.reshape(T, 9)
    return H


def project_psd(H) -> np.ndarray:
    w, Q = np.linalg.eigh(0.5 * (H + np.swapaxes(H, 1, 2)))
    return np.einsum("tik,tk,tjk->tij", Q, np.maximum(w, 0.0), Q)


def elem_hessian_psd(pot, u) -> np.ndarray:
    """(T, 12, 12) PSD-projected element Hessians (see the module docstring)."""
    if isinstance(pot, ofem.Arap):
        return elem_hessian(pot, u)
    _, dhdX = _kinematics(pot, u)
    B = _B(dhdX)
    Hp = project_psd(snh_hessian_F(pot, u))
    return pot.dV[:, None, None] * np.einsum("tki,tkl,tlj->tij", B, Hp, B)


def vertex_blocks(pot, u, n_points, psd=False) -> np.ndarray:
    """(V, 3, 3): the block diagonal of the assembled Hessian."""
    H = elem_hessian_psd(pot, u) if psd else elem_hessian(pot, u)
    T = H.shape[0]
    out = np.zeros((n_points, 3, 3))
    H = H.reshape(T, 4, 3, 4, 3)
    for a in range(4):
        np.add.at(out, pot.cells[:, a], H[:, a, :, a, :])
    return out


def hess_prod(pot, u, p, n_points, psd=False) -> np.ndarray:
    H = elem_hessian_psd(pot, u) if psd else elem_hessian(pot, u)
    pc = np.asarray(p)[pot.cells].reshape(-1, 12)
    r = np.einsum("tij,tj->ti", H, pc).reshape(-1, 4, 3)
    out = np.zeros((n_points, 3))
    np.add.at(out, pot.cells.reshape(-1), r.reshape(-1, 3))
    return out


def hess_quad(pot, u, p, psd=False) -> float:
    """sum over cells of max(p_cell^T H_e p_cell, 0) (the reference's per-cell clamp, ``_base.py:379-380``)."""
    H = elem_hessian_psd(pot, u) if psd else elem_hessian(pot, u)
    pc = np.asarray(p)[pot.cells].reshape(-1, 12)
    return float(np.maximum(np.einsum("ti,tij,tj->t", pc, H, pc), 0.0).sum())
