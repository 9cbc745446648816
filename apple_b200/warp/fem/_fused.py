from __future__ import annotations

import ctypes
from types import SimpleNamespace

import numpy as np
import torch

from apple_b200 import _lib

from ._arap import Arap
from ._base import WarpPotentialFem
from ._stable_neo_hookean import StableNeoHookean


class FusedSnhArap(WarpPotentialFem):
    """The SUM of a ``StableNeoHookean`` and an ``Arap`` potential that share their cells, evaluated in
    one pass (one read of the tile, one vertex gather, one slot reduction, one RED per vertex and field).

    The reference evaluates such a model as two launches per operator (``warp/model/_model.py:13-36``),
    each re-reading the mesh; SURVEY.md section 8a row M-1.  Results equal ``snh.op + arap.op`` with the
    reference's clamps applied per potential."""

    KIND = _lib.KIND_SNH_ARAP
    MATERIAL_NAMES = ("lambda_", "mu")

    def __init__(self, snh: StableNeoHookean, arap: Arap, *, name: str | None = None):
        if not can_fuse(snh, arap):
            raise ValueError("potentials must share cells, dhdX, dtype and device to be fused")
        self.arap_materials = SimpleNamespace(mu=arap.materials.mu)
        self.arap_dV = arap.region.dV
        super().__init__(snh.region, snh.materials, points=snh.points, n_points=snh.n_points, dtype=snh.dtype,
                         device=snh.device, name=name or f"{snh.name}+{arap.name}", scatter=snh.scatter)
        self.parts = (snh.name, arap.name)

    def _create_handle(self) -> ctypes.c_void_p:
        handle = ctypes.c_void_p()
        n_cells = self.region.cells.shape[0]
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        npdt = _lib.np_dtype(self.dtype)
        dv2 = np.ascontiguousarray(self.arap_dV, dtype=npdt).reshape(n_cells)
        mu2 = np.ascontiguousarray(self.arap_materials.mu, dtype=npdt)
        _lib.check(
            _lib.lib().apl_fem_create_snh_arap(
                _lib.dtype_code(self.dtype), n_cells, self.n_points, _lib.host_ptr(self.region.cells),
                _lib.host_ptr(self.region.dhdX.reshape(n_cells, 4, 3)), _lib.host_ptr(self.region.dV.reshape(n_cells)),
                _lib.host_ptr(self.materials.mu), _lib.host_ptr(self.materials.lambda_), _lib.host_ptr(dv2),
                _lib.host_ptr(mu2), _lib.host_ptr(self.points), dev_index, ctypes.byref(handle),
            )
        )
        return handle

    @classmethod
    def from_device_mesh(cls, cells, points, *, mu, lambda_, mu_arap, fraction=None, fraction_arap=None, **kw):
        """Fused pair built on the GPU (see ``WarpPotentialFem.from_device_mesh``): ``mu`` / ``lambda_`` / ``fraction`` are
        the Stable Neo-Hookean half, ``mu_arap`` / ``fraction_arap`` the ARAP half."""
        kw.setdefault("name", "snh+arap")
        second = {"mu": mu_arap}
        if fraction_arap is not None:
            second["fraction"] = fraction_arap
        self = super().from_device_mesh(cells, points, fraction=fraction, mu=mu, lambda_=lambda_, _second=second, **kw)
        self.arap_materials = SimpleNamespace(mu=mu_arap)
        self.arap_dV = None
        self.parts = ("snh", "arap")
        return self

    def mixed_derivative_prod(self, u, p) -> dict:
        raise NotImplementedError("evaluate mixed_derivative_prod on the two potentials of the fused pair separately")

    def set_materials(self, **kw) -> None:
        raise NotImplementedError("rebuild the fused potential after changing the materials of its parts")


def can_fuse(a, b) -> bool:
    return (
        type(a) is StableNeoHookean and type(b) is Arap and a.dtype == b.dtype and a.device == b.device
        and a.n_points == b.n_points and a.region.cells.shape == b.region.cells.shape
        and (a.region.cells is b.region.cells or np.array_equal(a.region.cells, b.region.cells))
        and (a.region.dhdX is b.region.dhdX or np.array_equal(a.region.dhdX, b.region.dhdX))
    )


def fuse_potentials(potentials: dict) -> dict:
    """Replaces every (StableNeoHookean, Arap) pair over identical cells by one ``FusedSnhArap``."""
    out, used = {}, set()
    items = list(potentials.items())
    for i, (na, a) in enumerate(items):
        if na in used:
            continue
        partner = None
        if type(a) in (StableNeoHookean, Arap):
            for nb, b in items[i + 1:]:
                if nb in used:
                    continue
                snh, arap = (a, b) if type(a) is StableNeoHookean else (b, a)
                if can_fuse(snh, arap):
                    partner = (nb, snh, arap)
                    break
        if partner is None:
            out[na] = a
        else:
            nb, snh, arap = partner
            used.add(nb)
            fused = FusedSnhArap(snh, arap)
            out[fused.name] = fused
    return out
