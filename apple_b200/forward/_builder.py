from __future__ import annotations

from apple_b200.warp.model import WarpModel, WarpModelAdapter, WarpPotential

from ._model import Model
from .dof_map import DofMapBuilder


class ModelBuilder:
    """Mirror of ``forward/_builder.py:12-40``."""

    def __init__(self, *, dtype=None, device=None):
        self.collision = None
        self.dof = DofMapBuilder()
        self.potentials: list[WarpPotential] = []
        self.dtype = dtype
        self.device = device

    def add_fixed(self, obj) -> None:
        self.dof.add_fixed(obj)

    def add_potential(self, potential: WarpPotential) -> None:
        self.potentials.append(potential)

    def add_vertices(self, obj) -> None:
        self.dof.add_vertices(obj)

    def finalize(self, *, fuse: bool = True) -> Model:
        """``fuse=True`` evaluates potentials that share their cells in one pass (same results as the
        reference's sum over potentials, ``warp/model/_model.py:13-36``)."""
        collision = None
        if self.collision is not None:
            collision = self.collision.finalize()
        dtype, device = self.dtype, self.device
        if self.potentials and dtype is None:
            dtype = getattr(self.potentials[0], "dtype", None)
        if self.potentials and device is None:
            device = getattr(self.potentials[0], "device", None)
        dof_map = self.dof.finalize(dtype=dtype, device=device)
        potentials = {potential.name: potential for potential in self.potentials}
        if fuse:
            from apple_b200.warp.fem import fuse_potentials

            potentials = fuse_potentials(potentials)
        warp_model = WarpModel(potentials)
        adapter = WarpModelAdapter(warp_model, n_points=dof_map.n_points)
        return Model(dof_map=dof_map, warp_model=adapter, collision=collision)
