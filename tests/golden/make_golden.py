"""Generates tests/golden/*.npz from the oracle (numpy fp64).

The reference itself cannot be imported in this environment (warp / jax / pyvista are absent), so
these fixtures are outputs of the ORACLE -- which is pinned by the reference's known-answer test and
by finite differences (tests/test_oracle.py) -- on small seeded meshes.  They freeze the oracle's
behaviour (regression) and give the GPU tests inputs/outputs that do not depend on regenerating
anything at run time.  Run from the repo root:  python tests/golden/make_golden.py
"""

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from helpers import KINDS, make_case, oracle_potential  # noqa: E402
from oracle import fem as ofem, pncg as opncg, region as oregion  # noqa: E402


def operators():
    mesh, u, p = make_case(n=3, seed=21)
    out = {"points": mesh.points, "cells": mesh.cells, "u": u, "p": p}
    for k, v in mesh.cell_data.items():
        out["cell_" + k] = v
    V = mesh.n_points
    for kind in KINDS:
        pot = oracle_potential(kind, mesh)
        m = ofem.Model([pot], V)
        out[f"{kind}_fun"] = m.fun(u)
        out[f"{kind}_grad"] = m.grad(u)
        out[f"{kind}_hess_diag"] = m.hess_diag(u)
        out[f"{kind}_hess_prod"] = m.hess_prod(u, p)
        out[f"{kind}_hess_quad"] = m.hess_quad(u, p)
        out[f"{kind}_elem_fun"] = pot.elem_fun(u)
        out[f"{kind}_elem_grad"] = pot.elem_grad(u)
        out[f"{kind}_elem_hess_prod"] = pot.elem_hess_prod(u, p)
    np.savez_compressed(Path(__file__).parent / "operators_n3.npz", **out)


def kat():
    """The reference's known-answer test: tests/forward/test_static_simulation.py:16-93."""
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0.25, 0.25, 0.25]], float)
    cells = np.array([[0, 1, 2, 4], [0, 1, 4, 3], [0, 4, 2, 3], [4, 1, 2, 3]])
    dhdX, dV = oregion.compute_grad(pts, cells)
    mask = np.zeros((5, 3), bool); mask[:4] = True
    val = np.zeros((5, 3)); val[3] = [0.2, -0.1, 0.15]
    problem = opncg.ForwardProblem(ofem.Model([ofem.Arap(cells, dhdX, dV, mu=np.ones(4))], 5), oregion.DofMap(mask, val))
    hist = []
    x, info = opncg.minimize(problem, np.zeros(3), max_steps=12, history=hist)
    np.savez_compressed(
        Path(__file__).parent / "kat_arap.npz", points=pts, cells=cells, fixed_mask=mask, fixed_value=val,
        expected_u4=np.array([0.05, -0.025, 0.0375]), energy_initial=problem.fun(np.zeros(3)), energy_final=info["fun"],
        pncg_energy_history=np.array([h[1] for h in hist]), pncg_gnorm_history=np.array([h[2] for h in hist]), x12=x,
    )


def pncg_cube():
    """Config-1 style solve at a size the oracle finishes in seconds: displacements after 25 iterations."""
    from apple_b200.mesh import lumped_vertex_volume

    mesh, _, _ = make_case(n=4, seed=11, grading=1.0)
    mesh.cell_data.pop("Fraction")
    V = mesh.n_points
    fixed = np.zeros((V, 3), bool); fixed[mesh.points[:, 2] == 0.0] = True
    idx = np.flatnonzero(~fixed[:, 0])
    force = np.zeros((idx.size, 3)); force[:, 0] = 40.0 * lumped_vertex_volume(mesh)[idx]
    model = ofem.Model([oracle_potential("snh", mesh), ofem.ExternalForce(force, idx)], V)
    problem = opncg.ForwardProblem(model, oregion.DofMap(fixed, np.zeros((V, 3))))
    x, info = opncg.minimize(problem, np.zeros(problem.dof_map.n_free), max_steps=25)
    np.savez_compressed(
        Path(__file__).parent / "pncg_cube_n4.npz", points=mesh.points, cells=mesh.cells, mu=mesh.cell_data["mu"],
        lam=mesh.cell_data["lambda"], fixed_mask=fixed, force=force, force_index=idx,
        u25=problem.dof_map.to_full(x), fun25=info["fun"], n_accepted=info["n_accepted"],
    )


if __name__ == "__main__":
    operators()
    kat()
    pncg_cube()
    print("golden fixtures written")
