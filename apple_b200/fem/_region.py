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
        X = self.points[self.cells_global]  # (c, a, I)
        dXdr = np.einsum("caI,aJ->cIJ", X, _DHDR)
        det = np.linalg.det(dXdr)
        if np.any(det == 0.0):
            raise ValueError("degenerate tetrahedron (zero rest volume)")
        drdX = np.linalg.inv(dXdr)
        dV = det * _WEIGHT
        if np.any(dV <= 0):
            logger.warning("dV <= 0")
        dhdX = np.einsum("aI,cIJ->caJ", _DHDR, drdX)
        self.dXdr = dXdr[:, None]
        self.drdX = drdX[:, None]
        self.dV = dV[:, None]
        self.dhdX = dhdX[:, None]
