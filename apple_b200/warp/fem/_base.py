from __future__ import annotations

import ctypes
from collections.abc import Sequence
from types import SimpleNamespace

import numpy as np
import torch

from apple_b200 import _lib, config
from apple_b200.common import ACTIVATION, FRACTION, LAMBDA, MU
from apple_b200.fem import Region
from apple_b200.warp.model import WarpPotential


def get_fraction(region: Region) -> np.ndarray:
    """``warp/fem/utils/_material.py:19-23``: cell_data["Fraction"] or ones."""
    fraction = region.cell_data.get(FRACTION.vtk)
    if fraction is None:
        return np.ones(region.n_cells)
    return np.asarray(fraction, dtype=np.float64)


def get_mu(region: Region) -> np.ndarray:  # _material.py:30-31
    return np.asarray(region.cell_data[MU.vtk], dtype=np.float64)


def get_lambda(region: Region) -> np.ndarray:  # _material.py:26-27
    return np.asarray(region.cell_data[LAMBDA.vtk], dtype=np.float64)


def get_activation(region: Region) -> np.ndarray:  # _material.py:15-16
    return np.asarray(region.cell_data[str(ACTIVATION)], dtype=np.float64).reshape(-1, 6)


class WarpPotentialFem(WarpPotential):
    """FEM potential over linear tets, mirror of ``warp/fem/_base.py:39-194``.

    Holds the region arrays (``region.cells``, ``region.dhdX``, ``region.dV``) and ``materials`` in
    the reference's layout on the host, and a native handle (``apl_fem_t``) with the packed, tiled
    device copy the kernels read.  The five operators launch one CUDA kernel each; ``eval`` fuses any
    subset into a single pass."""

    KIND: int = -1
    MATERIAL_NAMES: tuple[str, ...] = ()

    def __init__(self, region: SimpleNamespace, materials: SimpleNamespace, *, points=None, n_points=None,
                 dtype: torch.dtype | None = None, device=None, name: str | None = None,
                 requires_grad: Sequence[str] = (), scatter: int | None = None):
        super().__init__(name=name, requires_grad=requires_grad)
        self.dtype = dtype or config.default_dtype
        self.device = torch.device(device if device is not None else config.default_device())
        self.scatter = scatter
        npdt = _lib.np_dtype(self.dtype)
        self.region = SimpleNamespace(
            cells=np.ascontiguousarray(region.cells, dtype=np.int32),
            dhdX=np.ascontiguousarray(region.dhdX, dtype=npdt),
            dV=np.ascontiguousarray(region.dV, dtype=npdt),
        )
        self.materials = SimpleNamespace(
            **{k: np.ascontiguousarray(getattr(materials, k), dtype=npdt) for k in self.MATERIAL_NAMES}
        )
        n_cells = self.region.cells.shape[0]
        self.n_cells = int(n_cells)
        self.n_points = int(n_points if n_points is not None else (self.region.cells.max() + 1 if n_cells else 1))
        if self.device.type != "cuda":
            raise _lib.NativeError("apple_b200 potentials live on a CUDA device (there is no CPU path)")
        self.points = None if points is None else np.ascontiguousarray(points, dtype=np.float64)
        self._handle = self._create_handle()

    def _create_handle(self) -> ctypes.c_void_p:
        handle = ctypes.c_void_p()
        n_cells = self.region.cells.shape[0]
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(
            _lib.lib().apl_fem_create(
                self.KIND, _lib.dtype_code(self.dtype), n_cells, self.n_points,
                _lib.host_ptr(self.region.cells), _lib.host_ptr(self.region.dhdX.reshape(n_cells, 4, 3)),
                _lib.host_ptr(self.region.dV.reshape(n_cells)),
                _lib.host_ptr(getattr(self.materials, "mu", None)),
                _lib.host_ptr(getattr(self.materials, "lambda_", None)),
                _lib.host_ptr(getattr(self.materials, "activation", None)),
                _lib.host_ptr(self.points), dev_index, ctypes.byref(handle),
            )
        )
        return handle

    def __del__(self):
        handle = getattr(self, "_handle", None)
        if handle is not None and handle.value:
            try:
                _lib.lib().apl_fem_destroy(handle)
            except Exception:  # interpreter shutdown
                pass
            self._handle = None

    # ---- constructors (``_base.py:76-111``) ----
    @classmethod
    def from_pyvista(cls, obj, **kwargs):
        region = Region.from_pyvista(obj, grad=True)
        return cls.from_region(region, **kwargs)

    @classmethod
    def from_region(cls, region: Region, *, requires_grad: Sequence[str] = (), **kwargs):
        fraction = get_fraction(region)
        region_np = SimpleNamespace(
            cells=region.cells_global,
            dhdX=region.dhdX,  # (cells, quadrature=1, 4, 3)
            dV=fraction[:, None] * region.dV,  # _base.py:104
        )
        kwargs.setdefault("points", region.points)
        kwargs.setdefault("n_points", region.points.shape[0])
        return cls(region_np, cls.materials_from_region(region, requires_grad), requires_grad=requires_grad, **kwargs)

    @classmethod
    def materials_from_region(cls, region: Region, requires_grad: Sequence[str]) -> SimpleNamespace:
        raise NotImplementedError

    @classmethod
    def from_device_mesh(cls, cells: torch.Tensor, points: torch.Tensor, *, fraction=None, dtype: torch.dtype | None = None,
                         name: str | None = None, requires_grad: Sequence[str] = (), scatter: int | None = None,
                         morton: bool = True, _second=None, **materials):
        """Setup ENTIRELY on the GPU (``apl_fem_create_from_mesh``): ``cells`` (n_cells, 4) int32 and ``points``
        (n_points, 3) float64 are CUDA tensors, the materials (``mu=``, ``lambda_=``, ``activation=``) and the optional
        ``fraction`` (``cell_data["Fraction"]``) CUDA tensors of ``dtype``.  The rest shape -- what
        ``Region.compute_grad`` (``jax/fem/region/_region.py:84-108``) and ``from_region`` (``warp/fem/_base.py:93-111``)
        produce on the host -- is computed by a kernel straight into the packed planes; nothing but the connectivity
        (for the tile tables) visits the host.  ``self.region`` then holds no host arrays."""
        self = cls.__new__(cls)
        WarpPotential.__init__(self, name=name, requires_grad=requires_grad)
        self.dtype = dtype or config.default_dtype
        if not (cells.is_cuda and points.is_cuda):
            raise _lib.NativeError("from_device_mesh takes CUDA tensors (there is no CPU path)")
        self.device = cells.device
        self.scatter = scatter
        if cells.dtype != torch.int32 or cells.dim() != 2 or cells.shape[1] != 4:
            raise TypeError("cells: expected an (n_cells, 4) int32 tensor")
        if points.dtype != torch.float64 or points.dim() != 2 or points.shape[1] != 3:
            raise TypeError("points: expected an (n_points, 3) float64 tensor")
        cells, points = cells.contiguous(), points.contiguous()
        n_cells = int(cells.shape[0])
        missing = [k for k in cls.MATERIAL_NAMES if k not in materials]
        if missing or set(materials) - set(cls.MATERIAL_NAMES):
            raise KeyError(f"{cls.__name__}.from_device_mesh needs exactly the materials {cls.MATERIAL_NAMES}")

        def dev(a, shape):
            if a is None:
                return None
            t = torch.as_tensor(a, dtype=self.dtype, device=self.device).contiguous()
            if tuple(t.shape) != shape:
                raise ValueError(f"expected a tensor of shape {shape}, got {tuple(t.shape)}")
            return t

        mats = {k: dev(v, (n_cells, 6) if k == "activation" else (n_cells,)) for k, v in materials.items()}
        fraction = dev(fraction, (n_cells,))
        second = {k: dev(v, (n_cells,)) for k, v in (_second or {}).items()}
        self.region = SimpleNamespace(cells=None, dhdX=None, dV=None)      # no host copies on this path
        self.materials = SimpleNamespace(**mats)
        self.n_cells, self.n_points, self.points = n_cells, int(points.shape[0]), None
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.current_stream(self.device).synchronize()      # the setup kernels run on the default stream
            _lib.check(_lib.lib().apl_fem_create_from_mesh(
                cls.KIND, _lib.dtype_code(self.dtype), n_cells, self.n_points, _lib.dev_ptr(cells), _lib.dev_ptr(points),
                _lib.dev_ptr(fraction), _lib.dev_ptr(mats.get("mu")), _lib.dev_ptr(mats.get("lambda_")),
                _lib.dev_ptr(mats.get("activation")), _lib.dev_ptr(second.get("fraction")), _lib.dev_ptr(second.get("mu")),
                1 if morton else 0, self.device.index if self.device.index is not None else torch.cuda.current_device(),
                ctypes.byref(handle)))
        self._handle = handle
        return self

    @property
    def launch_dim(self) -> tuple[int, int]:
        return (self.n_cells, 1)

    @property
    def info(self) -> dict:
        buf = (ctypes.c_int64 * 10)()
        _lib.check(_lib.lib().apl_fem_info(self._handle, buf))
        keys = ("n_cells", "n_points", "n_tiles", "n_tile_verts", "static_bytes", "kind", "dtype", "device",
                "n_tile_voff")
        return dict(zip(keys, list(buf)))

    def set_materials(self, *, dV=None, **materials) -> None:
        """Replace per-cell materials (and/or ``dV``) without rebuilding the tiling."""
        npdt = _lib.np_dtype(self.dtype)
        arrs = {}
        if self.region.cells is None:
            raise NotImplementedError("set_materials: rebuild a potential created by from_device_mesh instead")
        if dV is not None:
            self.region.dV = np.ascontiguousarray(dV, dtype=npdt).reshape(self.region.dV.shape)
            arrs["dV"] = self.region.dV
        for k, v in materials.items():
            if k not in self.MATERIAL_NAMES:
                raise KeyError(f"{type(self).__name__} has no material {k!r}")
            setattr(self.materials, k, np.ascontiguousarray(v, dtype=npdt))
            arrs[k] = getattr(self.materials, k)
        _lib.check(
            _lib.lib().apl_fem_set_materials(
                self._handle, _lib.host_ptr(arrs.get("dV")), _lib.host_ptr(arrs.get("mu")),
                _lib.host_ptr(arrs.get("lambda_")), _lib.host_ptr(arrs.get("activation")),
            )
        )

    def mark_boundary(self, vertex_flags) -> int:
        """Tiles touching a vertex with a non-zero flag become the BOUNDARY part (``eval(..., part=1)``), the
        others the INTERIOR part (``part=2``); returns the number of boundary tiles.  Used by the sharded
        operators to overlap the halo exchange with the interior pass (``apple_b200/dist/_model.py``)."""
        n = ctypes.c_int64()
        flags = None if vertex_flags is None else np.ascontiguousarray(vertex_flags, dtype=np.uint8)
        if flags is not None and flags.size != self.n_points:
            raise ValueError("vertex_flags must have one entry per point")
        _lib.check(_lib.lib().apl_fem_mark_boundary(self._handle, _lib.host_ptr(flags), ctypes.byref(n)))
        return int(n.value)

    # ---- operators ----
    def eval(self, ops: int, u, p=None, *, fun=None, quad=None, grad=None, diag=None, prod=None, offd=None, scatter=None,
             part: int = 0) -> None:
        """Any OR of the operator bits in ONE pass.  ``offd`` (``OP_HESS_OFFD``, opt-in): the off-diagonal entries
        (xy, xz, yz) of the 3x3 vertex blocks; it travels through the ``prod`` slot of the C ABI, so it excludes
        ``OP_HESS_PROD`` / ``OP_HESS_QUAD``.  ``OP_PSD`` (opt-in) switches the Hessian terms to the eigenvalue-clamped
        ``d2Psi/dF2``."""
        if ops & _lib.OP_HESS_OFFD:
            if ops & (_lib.OP_HESS_PROD | _lib.OP_HESS_QUAD) or prod is not None:
                raise ValueError("OP_HESS_OFFD combines with OP_FUN, OP_GRAD and OP_HESS_DIAG only")
            prod = offd
        ld_in = _lib.field_ld(u, self.n_points, self.dtype, "u")
        if ops & (_lib.OP_HESS_PROD | _lib.OP_HESS_QUAD):
            if p is None or _lib.field_ld(p, self.n_points, self.dtype, "p") != ld_in:
                raise ValueError("p must be given with the same layout as u for hess_prod / hess_quad")
        ld_out = None
        for name, out, bit in (("grad", grad, _lib.OP_GRAD), ("diag", diag, _lib.OP_HESS_DIAG),
                               ("prod", prod, _lib.OP_HESS_PROD | _lib.OP_HESS_OFFD)):
            if ops & bit:
                ld = _lib.field_ld(out, self.n_points, self.dtype, name)
                if ld_out is not None and ld != ld_out:
                    raise ValueError("all output fields of one call must share a layout")
                ld_out = ld
        for name, out, bit in (("fun", fun, _lib.OP_FUN), ("quad", quad, _lib.OP_HESS_QUAD)):
            if ops & bit and (out is None or out.dtype != self.dtype or out.numel() < 1):
                raise ValueError(f"{name}: expected a ({self.dtype}) tensor with one element")
        if scatter is None:
            scatter = self.scatter if self.scatter is not None else config.scatter
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.lib().apl_fem_eval_part(
                    self._handle, int(part), ops, _lib.dev_ptr(u), _lib.dev_ptr(p), ld_in, _lib.dev_ptr(fun), _lib.dev_ptr(quad),
                    _lib.dev_ptr(grad), _lib.dev_ptr(diag), _lib.dev_ptr(prod), ld_out or 3, scatter,
                    _lib.stream_ptr(self.device),
                )
            )

    def mixed_derivative_prod(self, u, p) -> dict:
        """``d/dq [grad_u E(u) . p]`` per cell for every material ``q`` of this potential (caller's cell order): what
        the reference's inverse problems add to ``dL/dq`` after the adjoint solve
        (``exp/2026/01/28/smas/src/31-inverse-activation-stable-neo-hookean.py:472-487``)."""
        ld_in = _lib.field_ld(u, self.n_points, self.dtype, "u")
        if _lib.field_ld(p, self.n_points, self.dtype, "p") != ld_in:
            raise ValueError("p must have the layout of u")
        n_cells = self.n_cells
        out = {}
        for name in self.MATERIAL_NAMES:
            shape = (n_cells, 6) if name == "activation" else (n_cells,)
            out[name] = torch.zeros(shape, dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().apl_fem_mixed_derivative_prod(
                self._handle, _lib.dev_ptr(u), _lib.dev_ptr(p), ld_in, _lib.dev_ptr(out.get("mu")),
                _lib.dev_ptr(out.get("lambda_")), _lib.dev_ptr(out.get("activation")), _lib.stream_ptr(self.device)))
        return out

    def hess_block(self, u, diag, offd, *, psd: bool = False) -> None:
        """Opt-in superset of ``hess_diag``: the 3x3 vertex blocks of the assembled Hessian, accumulated as their
        diagonals ``diag[v] = (xx, yy, zz)`` -- the same numbers ``hess_diag`` returns -- and off-diagonals
        ``offd[v] = (xy, xz, yz)``.  ``psd=True``: blocks of the eigenvalue-clamped element Hessians."""
        self.eval(_lib.OP_HESS_DIAG | _lib.OP_HESS_OFFD | (_lib.OP_PSD if psd else 0), u, diag=diag, offd=offd)

    def hess_prod_psd(self, u, p, output) -> None:
        """``hess_prod`` with the analytic PSD projection of every element Hessian (opt-in; a descent-safe product)."""
        self.eval(_lib.OP_HESS_PROD | _lib.OP_PSD, u, p, prod=output)

    def hess_quad_psd(self, u, p, output) -> None:
        self.eval(_lib.OP_HESS_QUAD | _lib.OP_PSD, u, p, quad=output)

    def fun(self, u, output) -> None:  # _base.py:151-158
        self.eval(_lib.OP_FUN, u, fun=output)

    def grad(self, u, output) -> None:  # _base.py:160-167
        self.eval(_lib.OP_GRAD, u, grad=output)

    def hess_diag(self, u, output) -> None:  # _base.py:169-176
        self.eval(_lib.OP_HESS_DIAG, u, diag=output)

    def hess_prod(self, u, p, output) -> None:  # _base.py:178-185
        self.eval(_lib.OP_HESS_PROD, u, p, prod=output)

    def hess_quad(self, u, p, output) -> None:  # _base.py:187-194
        self.eval(_lib.OP_HESS_QUAD, u, p, quad=output)
