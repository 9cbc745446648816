"""Oracle: rest-shape precompute and DOF maps (test infrastructure, see ``oracle/__init__``).

Restates
* ``jax/fem/element/_tetra.py:36-45`` (reference-element shape-function gradients),
* ``jax/fem/quadrature/_tetra.py:12-15`` (one-point rule, weight 1/6),
* ``jax/fem/region/_region.py:84-108`` (``Region.compute_grad``),
* ``warp/fem/_base.py:93-111`` + ``warp/fem/utils/_material.py:19-23`` (``dV *= Fraction``),
* ``forward/dof_map/_builder.py:52-64`` and ``forward/dof_map/_dof_map.py:30-49`` (DOF maps).
"""

from __future__ import annotations

import numpy as np

# jax/fem/element/_tetra.py:36-45 -- d h_a / d r_J for the linear tetrahedron
DHDR = np.array(
    [[-1.0, -1.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
)
# jax/fem/quadrature/_tetra.py:14-15
QUAD_WEIGHT = 1.0 / 6.0


def compute_grad(points: np.ndarray, cells: np.ndarray, fraction=None, dtype=np.float64):
    """``Region.compute_grad`` (region/_region.py:84-108) for linear tets, q = 1.

    Returns ``dhdX (T,4,3)`` and ``dV (T,)`` (already multiplied by ``Fraction`` as
    ``WarpPotentialFem.from_region`` does, _base.py:104).
    """
    points = np.asarray(points, dtype=np.float64)
    cells = np.asarray(cells)
    X = points[cells]  # (T,4,3)  "c a I"
    # dXdr[c,I,J] = sum_a X[c,a,I] dhdr[a,J]            (_region.py:91-93)
    dXdr = np.einsum("caI,aJ->cIJ", X, DHDR)
    drdX = np.linalg.inv(dXdr)  # (_region.py:94)
    dV = np.linalg.det(dXdr) * QUAD_WEIGHT  # (_region.py:95-97)
    # dhdX[c,a,J] = sum_I dhdr[a,I] drdX[c,I,J]          (_region.py:100-102)
    dhdX = np.einsum("aI,cIJ->caJ", DHDR, drdX)
    if fraction is not None:
        dV = np.asarray(fraction, dtype=np.float64) * dV
    return dhdX.astype(dtype), dV.astype(dtype)


class DofMap:
    """``DofMap`` (dof_map/_dof_map.py:10-49) built as ``DofMapBuilder.finalize`` does
    (dof_map/_builder.py:52-64): per-component mask flattened row-major."""

    def __init__(self, fixed_mask: np.ndarray, fixed_value: np.ndarray):
        fixed_mask = np.asarray(fixed_mask, dtype=bool)
        self.n_points, self.dim = fixed_mask.shape
        self.fixed_indices = np.flatnonzero(fixed_mask)
        self.fixed_values = np.asarray(fixed_value).reshape(-1)[self.fixed_indices]
        self.free_indices = np.flatnonzero(~fixed_mask)

    @property
    def n_free(self) -> int:
        return self.free_indices.size

    def to_free(self, full: np.ndarray) -> np.ndarray:  # _dof_map.py:30-37
        return full.reshape(-1)[self.free_indices]

    def to_full(self, free: np.ndarray) -> np.ndarray:  # _dof_map.py:39-43
        out = np.empty(self.n_points * self.dim, dtype=free.dtype)
        out[self.fixed_indices] = self.fixed_values
        out[self.free_indices] = free
        return out.reshape(self.n_points, self.dim)

    def to_full_grad(self, free: np.ndarray) -> np.ndarray:  # _dof_map.py:45-49
        out = np.zeros(self.n_points * self.dim, dtype=free.dtype)
        out[self.free_indices] = free
        return out.reshape(self.n_points, self.dim)
