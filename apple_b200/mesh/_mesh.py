from __future__ import annotations

import numpy as np

VTK_TETRA = 10


class TetMesh:
    """Minimal stand-in for ``pyvista.UnstructuredGrid`` restricted to linear tetrahedra."""

    def __init__(self, points, cells, point_data=None, cell_data=None):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        cells = np.asarray(cells)
        if cells.ndim == 1:  # VTK flat layout [4, a, b, c, d, 4, ...] as in the reference's test mesh
            cells = cells.reshape(-1, 5)
            if not np.all(cells[:, 0] == 4):
                raise ValueError("flat cell array must describe tetrahedra (leading 4s)")
            cells = cells[:, 1:]
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        if self.cells.ndim != 2 or self.cells.shape[1] != 4:
            raise ValueError("cells must have shape (n_cells, 4)")
        self.point_data: dict = dict(point_data or {})
        self.cell_data: dict = dict(cell_data or {})

    @property
    def n_points(self) -> int:
        return self.points.shape[0]

    @property
    def n_cells(self) -> int:
        return self.cells.shape[0]

    @property
    def cells_dict(self) -> dict:
        return {VTK_TETRA: self.cells}

    @classmethod
    def load(cls, path) -> "TetMesh":
        """``pv.read(path)`` for ``.vtu`` files of linear tetrahedra (``apple_b200/mesh/_vtu.py``)."""
        from ._vtu import read_vtu

        return read_vtu(path)

    def save(self, path, *, binary: bool = True, compress: bool = False) -> None:
        """``mesh.save(path)`` of pyvista: VTK XML UnstructuredGrid with every point / cell data array."""
        from ._vtu import write_vtu

        write_vtu(self, path, binary=binary, compress=compress)

    def copy(self) -> "TetMesh":
        return TetMesh(
            self.points.copy(),
            self.cells.copy(),
            {k: np.array(v, copy=True) for k, v in self.point_data.items()},
            {k: np.array(v, copy=True) for k, v in self.cell_data.items()},
        )


def as_tet_arrays(obj):
    """(points float64 (V,3), cells int32 (T,4)) of a ``TetMesh`` or a pyvista unstructured grid."""
    if isinstance(obj, TetMesh):
        return obj.points, obj.cells
    cells_dict = getattr(obj, "cells_dict", None)
    if cells_dict is None or VTK_TETRA not in cells_dict:
        raise TypeError("expected a TetMesh or a pyvista.UnstructuredGrid made of tetrahedra")
    return (
        np.ascontiguousarray(obj.points, dtype=np.float64),
        np.ascontiguousarray(cells_dict[VTK_TETRA], dtype=np.int32),
    )
