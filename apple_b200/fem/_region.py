"""``Region``: one-time rest-shape precompute for linear tetrahedra.

Mirrors ``jax/fem/region/_region.py:18-129`` (``Region.from_pyvista(mesh, grad=True)``,
``cells_global``, ``dhdX (c,q,a,J)``, ``dV (c,q)``, ``cell_data``, ``point_data``) with the
linear-tet element of ``jax/fem/element/_tetra.py:36-45`` and the one-point rule of
``jax/fem/quadrature/_tetra.py:12-15``.  Setup-time host code (numpy float64); its OUTPUT LAYOUT is
the input contract of ``apl_fem_create``.
"""

from __future__ import annotations

import logging

import numpy as np

from apple_b200.mesh import as_tet_arrays

logger = logging.getLogger(__name__)

_DHDR = np.array([[-1.0, -1.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
_WEIGHT = 1.0 / 6.0


class Region:
    def __init__(self, mesh):
        self.mesh = mesh
        self.points, self.cells_global = as_tet_arrays(mesh)
        self.dhdr = _DHDR[None]  # (q, a, J)
        self.dXdr = self.drdX = self.dV = self.dhdX = None

    @classmethod
    def from_pyvista(cls, mesh, *, grad: bool = False, quadrature=None) -> "Region":
        if quadrature is not None:
            raise NotImplementedError("only the default one-point tetrahedral rule is supported")
        self = cls(mesh)
        if grad:
            self.compute_grad()
        return self

    @property
    def n_cells(self) -> int:
        return self.cells_global.shape[0]

    @property
    def cells_local(self) -> np.ndarray:
        return self.cells_global

    @property
    def point_data(self):
        return self.mesh.point_data

    @property
    def cell_data(self):
        return self.mesh.cell_data

    def compute_grad(self) -> None:
        """Closed form of ``Region.compute_grad`` for the linear tetrahedron: ``dXdr`` has the edge vectors
        ``X_a - X_0`` as columns, ``drdX`` is its inverse written with cross products, ``dV = det / 6`` and
        ``dhdX = dhdr . drdX`` (rows 1..3 are the rows of the inverse, row 0 minus their sum)."""
        X = self.points[self.cells_global]  # (c, a, I)
        e1, e2, e3 = X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]
        c23, c31, c12 = np.cross(e2, e3), np.cross(e3, e1), np.cross(e1, e2)
        det = np.einsum("ci,ci->c", e1, c23)
        if np.any(det == 0.0):
            raise ValueError("degenerate tetrahedron (zero rest volume)")
        inv = np.stack([c23, c31, c12], axis=1) / det[:, None, None]  # rows of (dXdr)^-1
        dV = det * _WEIGHT
        if np.any(dV <= 0):
            logger.warning("dV <= 0")
        dhdX = np.concatenate([-inv.sum(axis=1, keepdims=True), inv], axis=1)
        self.dXdr = np.stack([e1, e2, e3], axis=2)[:, None]
        self.drdX = inv[:, None]
        self.dV = dV[:, None]
        self.dhdX = dhdX[:, None]
