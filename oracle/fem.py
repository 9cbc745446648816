"""Oracle: per-tetrahedron energies and their assembly (test infrastructure).

A literal numpy restatement, vectorised over cells, of

* ``warp/fem/func/_deformation.py:9-31``  (F, jvp, vjp)
* ``warp/fem/func/_identity.py:20-39``    (I2, I3)
* ``warp/fem/func/_gradient.py:19-40``    (g2, g3)
* ``warp/fem/func/_hess_diag.py:29-72``   (h3, h4, h5, h6 diagonals)
* ``warp/fem/func/_hess_prod.py:29-77``   (h3, h4, h5, h6 products)
* ``warp/fem/func/_hess_quad.py:30-74``   (h3, h4, h5, h6 quadratic forms)
* ``warp/fem/func/_misc.py:22-70``        (gather, lambdas, activation matrix, Qs)
* ``warp/math/_rotation.py:9-22``         (svd_rv / polar_rv; ``wp.svd3`` -> numpy SVD made
  rotation-variant: U, V in SO(3), the smallest singular value may be negative)
* ``warp/fem/_stable_neo_hookean.py:17-103``, ``_stable_neo_hookean_muscle.py:18-114``,
  ``_arap.py:17-76``                      (the three energies)
* ``warp/fem/_base.py:243-383``           (the five assembly kernels, incl. both clamps)
* ``warp/potential/_ext_force.py:17-90``  (external force)
* ``warp/model/_model.py:9-36``           (zero, then sum over potentials)

Matrices are row-major numpy arrays with a leading cell axis: ``F (T,3,3)``, ``dhdX (T,4,3)``,
``u_cell / p_cell (T,4,3)``.
"""

from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------- func/_deformation


def gather(u: np.ndarray, cells: np.ndarray) -> np.ndarray:
    """``get_cell_displacements`` (func/_misc.py:22-28): (V,3),(T,4) -> (T,4,3)."""
    return u[cells]


def deformation_gradient(u_cell, dhdX):
    """``F = u^T dhdX + I`` (func/_deformation.py:15-19)."""
    return np.einsum("cai,caJ->ciJ", u_cell, dhdX) + np.eye(3, dtype=dhdX.dtype)


def jvp(dhdX, p_cell):
    """``deformation_gradient_jvp`` = p^T dhdX (func/_deformation.py:22-25)."""
    return np.einsum("cai,caJ->ciJ", p_cell, dhdX)


def vjp(dhdX, M):
    """``deformation_gradient_vjp`` = dhdX M^T (func/_deformation.py:28-31): (T,4,3)."""
    return np.einsum("caJ,ciJ->cai", dhdX, M)


def ddot(A, B):
    """``wp.ddot``: sum_ij A_ij B_ij per cell."""
    return (A * B).reshape(A.shape[0], -1).sum(axis=1)


# --------------------------------------------------------------------------- invariants


def I2(F):  # func/_identity.py:20-28
    return ddot(F, F)


def I3(F):  # func/_identity.py:31-39
    return np.linalg.det(F)


def g2(F):  # func/_gradient.py:19-27
    return 2.0 * F


def g3(F):
    """cof(F) with columns (f1 x f2, f2 x f0, f0 x f1) (func/_gradient.py:30-40)."""
    f0, f1, f2 = F[:, :, 0], F[:, :, 1], F[:, :, 2]
    return np.stack([np.cross(f1, f2), np.cross(f2, f0), np.cross(f0, f1)], axis=2)


# --------------------------------------------------------------------------- SVD pieces


def svd_rv(F):
    """``svd_rv`` (math/_rotation.py:9-13).  ``wp.svd3`` returns U, sigma, V with U, V proper
    rotations; numpy returns orthogonal factors and sigma >= 0, so reflections are moved into
    the smallest singular value.  For det F > 0 every quantity the energies use is independent
    of the remaining freedom."""
    U, s, Vh = np.linalg.svd(F)
    V = np.swapaxes(Vh, 1, 2).copy()
    U = U.copy()
    s = s.copy()
    neg_u = np.linalg.det(U) < 0
    U[neg_u, :, 2] *= -1.0
    s[neg_u, 2] *= -1.0
    neg_v = np.linalg.det(V) < 0
    V[neg_v, :, 2] *= -1.0
    s[neg_v, 2] *= -1.0
    return U, s, V


def polar_rv(F):
    """``polar_rv`` (math/_rotation.py:16-22): R = U V^T."""
    U, s, V = svd_rv(F)
    R = np.einsum("cij,ckj->cik", U, V)
    return R, (U, s, V)


def lambdas(sigma, clamp: bool = True):
    """func/_misc.py:31-43."""
    two = sigma.dtype.type(2.0)
    s01 = sigma[:, 0] + sigma[:, 1]
    s12 = sigma[:, 1] + sigma[:, 2]
    s20 = sigma[:, 2] + sigma[:, 0]
    if clamp:
        s01, s12, s20 = (np.maximum(x, two) for x in (s01, s12, s20))
    return np.stack([two / s01, two / s12, two / s20], axis=1)


def Qs(U, V):
    """Twist eigen-matrices (func/_misc.py:56-70)."""
    r = 1.0 / np.sqrt(U.dtype.type(2.0))
    outer = lambda a, b: np.einsum("ci,cj->cij", a, b)  # noqa: E731
    U0, U1, U2 = U[:, :, 0], U[:, :, 1], U[:, :, 2]
    V0, V1, V2 = V[:, :, 0], V[:, :, 1], V[:, :, 2]
    Q0 = (outer(U1, V0) - outer(U0, V1)) * r
    Q1 = (outer(U1, V2) - outer(U2, V1)) * r
    Q2 = (outer(U0, V2) - outer(U2, V0)) * r
    return Q0, Q1, Q2


def make_activation_mat33(a):
    """func/_misc.py:46-53: A = I + sym(a), a = (xx, yy, zz, xy, xz, yz)."""
    T = a.shape[0]
    A = np.empty((T, 3, 3), dtype=a.dtype)
    A[:, 0, 0] = 1.0 + a[:, 0]
    A[:, 1, 1] = 1.0 + a[:, 1]
    A[:, 2, 2] = 1.0 + a[:, 2]
    A[:, 0, 1] = A[:, 1, 0] = a[:, 3]
    A[:, 0, 2] = A[:, 2, 0] = a[:, 4]
    A[:, 1, 2] = A[:, 2, 1] = a[:, 5]
    return A


# --------------------------------------------------------------------------- h_k terms


def h3_diag(dhdX, g3_):  # func/_hess_diag.py:29-33
    return vjp(dhdX, g3_) ** 2


def h4_diag(dhdX, U, s, V, clamp_lambda=True):  # func/_hess_diag.py:36-50
    lam = lambdas(s, clamp_lambda)
    out = 0.0
    for i, Q in enumerate(Qs(U, V)):
        out = out + lam[:, i, None, None] * vjp(dhdX, Q) ** 2
    return out


def h5_diag(dhdX):  # func/_hess_diag.py:53-66
    t = np.einsum("caJ,caJ->ca", dhdX, dhdX)
    return 2.0 * np.repeat(t[:, :, None], 3, axis=2)


def h6_diag(dhdX, F):  # func/_hess_diag.py:69-72 (identically zero)
    return np.zeros_like(dhdX)


def h3_prod(p, dhdX, g3_):  # func/_hess_prod.py:29-34
    W = vjp(dhdX, g3_)
    return ddot(W, p)[:, None, None] * W


def h4_prod(p, dhdX, U, s, V, clamp_lambda=True):  # func/_hess_prod.py:37-52
    lam = lambdas(s, clamp_lambda)
    out = 0.0
    for i, Q in enumerate(Qs(U, V)):
        W = vjp(dhdX, Q)
        out = out + (lam[:, i] * ddot(W, p))[:, None, None] * W
    return out


def h5_prod(p, dhdX):  # func/_hess_prod.py:55-59
    return 2.0 * vjp(dhdX, jvp(dhdX, p))


def _h6_matrix(F, dF):
    f0, f1, f2 = F[:, :, 0], F[:, :, 1], F[:, :, 2]
    p0, p1, p2 = dF[:, :, 0], dF[:, :, 1], dF[:, :, 2]
    return np.stack(
        [
            np.cross(f1, p2) - np.cross(f2, p1),
            np.cross(f2, p0) - np.cross(f0, p2),
            np.cross(f0, p1) - np.cross(f1, p0),
        ],
        axis=2,
    )


def h6_prod(p, dhdX, F):  # func/_hess_prod.py:62-77
    return vjp(dhdX, _h6_matrix(F, jvp(dhdX, p)))


def h3_quad(p, dhdX, g3_):  # func/_hess_quad.py:30-34
    return ddot(jvp(dhdX, p), g3_) ** 2


def h4_quad(p, dhdX, U, s, V, clamp_lambda=True):  # func/_hess_quad.py:37-49
    dF = jvp(dhdX, p)
    lam = lambdas(s, clamp_lambda)
    out = 0.0
    for i, Q in enumerate(Qs(U, V)):
        out = out + lam[:, i] * ddot(Q, dF) ** 2
    return out


def h5_quad(p, dhdX):  # func/_hess_quad.py:52-56
    dF = jvp(dhdX, p)
    return 2.0 * ddot(dF, dF)


def h6_quad(p, dhdX, F):  # func/_hess_quad.py:59-74
    dF = jvp(dhdX, p)
    return ddot(dF, _h6_matrix(F, dF))


# --------------------------------------------------------------------------- potentials


class FemPotential:
    """``WarpPotentialFem`` (_base.py:39-194): region arrays + the five accumulate-into-output
    operators, restating the kernel factories at _base.py:243-383."""

    clamp_hess_diag = True  # _base.py:299
    clamp_hess_quad = True  # _base.py:359

    def __init__(self, cells, dhdX, dV, **materials):
        self.cells = np.asarray(cells)
        self.dhdX = np.asarray(dhdX)
        self.dV = np.asarray(dV)
        self.dtype = self.dhdX.dtype
        self.materials = {k: np.asarray(v, dtype=self.dtype) for k, v in materials.items()}

    # --- per-material device functions (overridden) ---
    def energy_density(self, F): ...
    def first_piola_kirchhoff(self, F): ...
    def hess_diag_func(self, F, dhdX): ...
    def hess_prod_func(self, F, p, dhdX): ...
    def hess_quad_func(self, F, p, dhdX): ...

    def _F(self, u):
        return deformation_gradient(gather(np.asarray(u, self.dtype), self.cells), self.dhdX)

    # --- per-element results (no scatter), used for element-level parity ---
    def elem_fun(self, u):  # _base.py:255-261
        return self.energy_density(self._F(u)) * self.dV

    def elem_grad(self, u):  # _base.py:277-287
        P = self.first_piola_kirchhoff(self._F(u))
        return vjp(self.dhdX, P) * self.dV[:, None, None]

    def elem_hess_diag(self, u):  # _base.py:307-320
        H = self.hess_diag_func(self._F(u), self.dhdX) * self.dV[:, None, None]
        if self.clamp_hess_diag:
            H = np.maximum(H, 0.0)
        return H

    def elem_hess_prod(self, u, p):  # _base.py:339-347
        p_cell = gather(np.asarray(p, self.dtype), self.cells)
        return self.dV[:, None, None] * self.hess_prod_func(self._F(u), p_cell, self.dhdX)

    def elem_hess_quad(self, u, p):  # _base.py:368-380
        p_cell = gather(np.asarray(p, self.dtype), self.cells)
        q = self.dV * self.hess_quad_func(self._F(u), p_cell, self.dhdX)
        if self.clamp_hess_quad:
            q = np.maximum(q, 0.0)
        return q

    # --- the five operators: accumulate into a caller-zeroed output (model/_potential.py) ---
    def fun(self, u, output):
        output[0] += self.elem_fun(u).sum()

    def _scatter(self, elem, output):  # the 4x atomic_add loops, e.g. _base.py:288-289
        np.add.at(output, self.cells.reshape(-1), elem.reshape(-1, 3))

    def grad(self, u, output):
        self._scatter(self.elem_grad(u), output)

    def hess_diag(self, u, output):
        self._scatter(self.elem_hess_diag(u), output)

    def hess_prod(self, u, p, output):
        self._scatter(self.elem_hess_prod(u, p), output)

    def hess_quad(self, u, p, output):
        output[0] += self.elem_hess_quad(u, p).sum()


class StableNeoHookean(FemPotential):
    """_stable_neo_hookean.py:17-103."""

    def _coeffs(self, J):
        mu, la = self.materials["mu"], self.materials["lambda_"]
        return 0.5 * mu, -mu + la * (J - 1.0), la  # dPsi_dI2, dPsi_dI3, d2Psi_dI32

    def energy_density(self, F):  # :17-27
        mu, la = self.materials["mu"], self.materials["lambda_"]
        J = I3(F)
        return 0.5 * mu * (I2(F) - 3.0) - mu * (J - 1.0) + 0.5 * la * (J - 1.0) ** 2

    def first_piola_kirchhoff(self, F):  # :30-39
        c2, c3, _ = self._coeffs(I3(F))
        return c2[:, None, None] * g2(F) + c3[:, None, None] * g3(F)

    def hess_diag_func(self, F, dhdX):  # :42-65
        c2, c3, c33 = self._coeffs(I3(F))
        b = lambda x: x[:, None, None]  # noqa: E731
        return b(c33) * h3_diag(dhdX, g3(F)) + b(c2) * h5_diag(dhdX) + b(c3) * h6_diag(dhdX, F)

    def hess_prod_func(self, F, p, dhdX):  # :68-83
        c2, c3, c33 = self._coeffs(I3(F))
        b = lambda x: x[:, None, None]  # noqa: E731
        return (
            b(c33) * h3_prod(p, dhdX, g3(F)) + b(c2) * h5_prod(p, dhdX) + b(c3) * h6_prod(p, dhdX, F)
        )

    def hess_quad_func(self, F, p, dhdX):  # :86-103
        c2, c3, c33 = self._coeffs(I3(F))
        return c33 * h3_quad(p, dhdX, g3(F)) + c2 * h5_quad(p, dhdX) + c3 * h6_quad(p, dhdX, F)


class StableNeoHookeanMuscle(StableNeoHookean):
    """_stable_neo_hookean_muscle.py:18-114: SNH evaluated on G = F A, dhdX -> dhdX A."""

    def _A(self):
        return make_activation_mat33(self.materials["activation"])

    def energy_density(self, F):  # :18-31
        return super().energy_density(F @ self._A())

    def first_piola_kirchhoff(self, F):  # :34-45
        A = self._A()
        return super().first_piola_kirchhoff(F @ A) @ np.swapaxes(A, 1, 2)

    def hess_diag_func(self, F, dhdX):  # :48-75
        A = self._A()
        return super().hess_diag_func(F @ A, dhdX @ A)

    def hess_prod_func(self, F, p, dhdX):  # :78-93
        A = self._A()
        return super().hess_prod_func(F @ A, p, dhdX @ A)

    def hess_quad_func(self, F, p, dhdX):  # :96-114
        A = self._A()
        return super().hess_quad_func(F @ A, p, dhdX @ A)


class Arap(FemPotential):
    """_arap.py:17-76.

    ``literal_reference_bug=True`` reproduces the reference's ``hess_prod`` argument swap
    (_arap.py:55-56 pass ``(dhdX, p)`` where ``func.h4_prod`` / ``func.h5_prod`` expect
    ``(p, dhdX)``).  That output is not the Hessian-vector product; it is kept here only to
    document the deviation.  The default is the mathematically correct product
    (SURVEY.md section 8a row A-ARAP)."""

    def __init__(self, *args, clamp_lambda=True, literal_reference_bug=False, **kw):
        super().__init__(*args, **kw)
        self.clamp_lambda = clamp_lambda
        self.literal_reference_bug = literal_reference_bug

    def energy_density(self, F):  # :17-21
        R, _ = polar_rv(F)
        D = F - R
        return 0.5 * self.materials["mu"] * ddot(D, D)

    def first_piola_kirchhoff(self, F):  # :24-28
        R, _ = polar_rv(F)
        return self.materials["mu"][:, None, None] * (F - R)

    def hess_diag_func(self, F, dhdX):  # :31-40
        U, s, V = svd_rv(F)
        h = -2.0 * h4_diag(dhdX, U, s, V, self.clamp_lambda) + h5_diag(dhdX)
        return 0.5 * self.materials["mu"][:, None, None] * h

    def hess_prod_func(self, F, p, dhdX):  # :43-58
        U, s, V = svd_rv(F)
        if self.literal_reference_bug:
            h = -2.0 * h4_prod(dhdX, p, U, s, V, self.clamp_lambda) + h5_prod(dhdX, p)
        else:
            h = -2.0 * h4_prod(p, dhdX, U, s, V, self.clamp_lambda) + h5_prod(p, dhdX)
        return 0.5 * self.materials["mu"][:, None, None] * h

    def hess_quad_func(self, F, p, dhdX):  # :61-76
        U, s, V = svd_rv(F)
        h = -2.0 * h4_quad(p, dhdX, U, s, V, self.clamp_lambda) + h5_quad(p, dhdX)
        return 0.5 * self.materials["mu"] * h


class ExternalForce:
    """potential/_ext_force.py:17-90: W = -sum_k f_k . u[idx_k]; zero Hessian."""

    def __init__(self, force, indices):
        self.force = np.asarray(force)
        self.indices = np.asarray(indices)

    def fun(self, u, output):  # :17-28
        output[0] += -(self.force * u[self.indices]).sum()

    def grad(self, u, output):  # :31-39
        np.add.at(output, self.indices, -self.force.astype(output.dtype))

    def hess_diag(self, u, output):  # :80-82
        pass

    def hess_prod(self, u, p, output):  # :84-86
        pass

    def hess_quad(self, u, p, output):  # :88-90
        pass


class Model:
    """``WarpModel`` (model/_model.py:9-36) + ``WarpModelAdapter`` return shapes
    (model/_adapter.py:21-39): zero the output, let every potential accumulate."""

    def __init__(self, potentials, n_points, dtype=np.float64):
        self.potentials = list(potentials)
        self.n_points = n_points
        self.dtype = np.dtype(dtype)

    def _scalar(self, name, *args):
        out = np.zeros(1, self.dtype)
        for pot in self.potentials:
            getattr(pot, name)(*args, out)
        return out[0]

    def _field(self, name, *args):
        out = np.zeros((self.n_points, 3), self.dtype)
        for pot in self.potentials:
            getattr(pot, name)(*args, out)
        return out

    def fun(self, u):
        return self._scalar("fun", u)

    def grad(self, u):
        return self._field("grad", u)

    def hess_diag(self, u):
        return self._field("hess_diag", u)

    def hess_prod(self, u, p):
        return self._field("hess_prod", u, p)

    def hess_quad(self, u, p):
        return self._scalar("hess_quad", u, p)
