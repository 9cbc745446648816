"""Shared builders for the parity tests: the same seeded mesh / materials / fields are handed to the
oracle (numpy) and to the CUDA path."""

from __future__ import annotations

import numpy as np

from apple_b200.mesh import TetMesh, cube_tet_mesh
from oracle import fem as ofem
from oracle import region as oregion

KINDS = ("snh", "arap", "muscle")


def make_case(n=6, seed=0, *, grading=1.02, heterogeneous=True, amp=0.05, morton=True):
    """Cube mesh with per-cell materials and smooth + random displacement / direction fields."""
    rng = np.random.default_rng(seed)
    mesh = cube_tet_mesh(n, grading=grading, morton=morton)
    T, V = mesh.n_cells, mesh.n_points
    if heterogeneous:
        mu = 10.0 ** rng.uniform(0.0, 2.0, T)
        la = 10.0 ** rng.uniform(0.0, 2.0, T)
    else:
        mu, la = np.full(T, 3.0), np.full(T, 7.0)
    mesh.cell_data["mu"] = mu
    mesh.cell_data["lambda"] = la
    mesh.cell_data["Fraction"] = rng.uniform(0.2, 1.0, T)
    act = 0.1 * rng.standard_normal((T, 6))
    mesh.cell_data["activation"] = act
    h = 1.0 / n
    X = mesh.points
    u = amp * h * (np.sin(3.0 * X[:, [1, 2, 0]] + 0.3) + 0.5 * rng.uniform(-1, 1, (V, 3)))
    p = rng.uniform(-1, 1, (V, 3))
    return mesh, np.ascontiguousarray(u), np.ascontiguousarray(p)


def oracle_potential(kind: str, mesh: TetMesh, dtype=np.float64):
    dhdX, dV = oregion.compute_grad(mesh.points, mesh.cells, mesh.cell_data.get("Fraction"), dtype=dtype)
    mu, la, act = mesh.cell_data["mu"], mesh.cell_data["lambda"], mesh.cell_data.get("activation")
    if kind == "snh":
        return ofem.StableNeoHookean(mesh.cells, dhdX, dV, mu=mu, lambda_=la)
    if kind == "arap":
        return ofem.Arap(mesh.cells, dhdX, dV, mu=mu)
    if kind == "muscle":
        return ofem.StableNeoHookeanMuscle(mesh.cells, dhdX, dV, mu=mu, lambda_=la, activation=act)
    raise KeyError(kind)


def cuda_potential(kind: str, mesh: TetMesh, dtype, **kw):
    from apple_b200.warp.fem import Arap, StableNeoHookean, StableNeoHookeanMuscle

    cls = {"snh": StableNeoHookean, "arap": Arap, "muscle": StableNeoHookeanMuscle}[kind]
    return cls.from_pyvista(mesh, dtype=dtype, **kw)


def rel_err(a, b) -> float:
    """max |a-b| / max |b| (fields), |a-b|/|b| (scalars)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.abs(b).max()
    return float(np.abs(a - b).max() / denom) if denom > 0 else float(np.abs(a - b).max())
