from __future__ import annotations

import numpy as np
import torch

from apple_b200 import config
from apple_b200.common import FIXED_MASK, FIXED_VALUE, GLOBAL_POINT_ID

from ._dof_map import DofMap


class DofMapBuilder:
    """Mirror of ``forward/dof_map/_builder.py:13-64``."""

    def __init__(self, dim: int = 3):
        self.dim = dim
        self.fixed_mask = np.empty((0, dim), dtype=bool)
        self.full_values = np.empty((0, dim), dtype=np.float64)

    @property
    def n_points(self) -> int:
        return self.fixed_mask.shape[0]

    def add_fixed(self, obj) -> None:
        gid = np.asarray(obj.point_data[GLOBAL_POINT_ID.vtk])
        self.fixed_mask[gid] = np.asarray(obj.point_data[FIXED_MASK.vtk], dtype=bool)
        self.full_values[gid] = np.asarray(obj.point_data[FIXED_VALUE.vtk], dtype=np.float64)

    def add_vertices(self, obj) -> None:
        obj.point_data[GLOBAL_POINT_ID.vtk] = np.arange(self.n_points, self.n_points + obj.n_points)
        self.fixed_mask = np.pad(self.fixed_mask, ((0, obj.n_points), (0, 0)), constant_values=False)
        self.full_values = np.pad(self.full_values, ((0, obj.n_points), (0, 0)), constant_values=0.0)

    def finalize(self, *, dtype=None, device=None) -> DofMap:
        dtype = dtype or config.default_dtype
        device = torch.device(device if device is not None else config.default_device())
        fixed_indices = np.flatnonzero(self.fixed_mask)
        fixed_values = self.full_values.reshape(-1)[fixed_indices]
        free_indices = np.flatnonzero(~self.fixed_mask)
        return DofMap(
            dim=self.dim,
            n_points=self.n_points,
            fixed_indices=torch.as_tensor(fixed_indices, dtype=torch.int64, device=device),
            fixed_values=torch.as_tensor(fixed_values, dtype=dtype, device=device),
            free_indices=torch.as_tensor(free_indices, dtype=torch.int64, device=device),
        )
