"""Synthetic meshes generated directly in device memory (torch ops; works on any torch device, so the CPU tests
compare it element for element with the numpy generators of ``_generate.py``).

A 64 M-tet cube is 25 GB of numpy arrays and minutes of host time; generated here it never leaves the GPU and feeds
``WarpPotentialFem.from_device_mesh`` (``apl_fem_create_from_mesh``), which orders, tiles and packs it on the device.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from ._generate import _EVEN, _ODD


def _oriented_templates():
    """The two 5-tet templates of ``_generate.py`` with every tet oriented to positive rest volume.  The sign of a
    template tet's volume does not depend on the (positive) grid spacings, so ``cube_tet_mesh``'s per-tet winding fix
    is a property of the template."""
    corner = np.array([[dx, dy, dz] for dx in (0, 1) for dy in (0, 1) for dz in (0, 1)], dtype=np.float64)
    out = []
    for tmpl in (_EVEN, _ODD):
        t = tmpl.copy()
        X = corner[t]
        vol = np.einsum("ci,ci->c", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 3] - X[:, 0])
        t[vol < 0] = t[vol < 0][:, [0, 2, 1, 3]]
        out.append(t)
    return out


_EVEN_POS, _ODD_POS = _oriented_templates()


def _line(m: int, grading: float, length: float, device) -> torch.Tensor:
    if grading == 1.0:
        return torch.linspace(0.0, length, m + 1, dtype=torch.float64, device=device)
    w = torch.tensor(grading, dtype=torch.float64, device=device) ** torch.arange(m, dtype=torch.float64, device=device)
    x = torch.cat([torch.zeros(1, dtype=torch.float64, device=device), torch.cumsum(w, 0)])
    return x / x[-1] * length


def morton_codes_device(xyz: torch.Tensor, lo: torch.Tensor, ext: float) -> torch.Tensor:
    """63-bit Morton code (int64) of points normalised to the cube ``lo + [0, ext]^3`` (``_generate.morton_codes``)."""
    def part1by2(x):
        x = x & 0x1FFFFF
        x = (x | (x << 32)) & 0x1F00000000FFFF
        x = (x | (x << 16)) & 0x1F0000FF0000FF
        x = (x | (x << 8)) & 0x100F00F00F00F00F
        x = (x | (x << 4)) & 0x10C30C30C30C30C3
        x = (x | (x << 2)) & 0x1249249249249249
        return x

    scale = ((1 << 21) - 1) / ext if ext > 0 else 0.0
    q = ((xyz - lo) * scale).to(torch.int64)
    return part1by2(q[:, 0]) | (part1by2(q[:, 1]) << 1) | (part1by2(q[:, 2]) << 2)


def cube_tet_slab_device(n: int, i0: int, i1: int, device, *, grading: float = 1.0, length: float = 1.0,
                         morton_vertices: bool = False):
    """The hex layers ``i0 <= i < i1`` (first grid axis) of the ``n``^3 x 5 cube, generated on ``device``: the same
    mesh as ``cube_tet_slab(n, i0, i1)`` (``i0 = 0, i1 = n``: ``cube_tet_mesh(n, morton=False)``), i.e.
    ``(points (V, 3) float64, cells (T, 4) int32 with local vertex ids, vertex_gid (V,) int64, cell_gid (T,) int64)``
    with the global ids ``(i (n+1) + j) (n+1) + k`` and ``5 hex + t`` of the full lexicographic mesh.
    ``morton_vertices=True`` numbers the LOCAL vertices along a Morton curve (the vertices a tile of Morton-ordered
    tets gathers are then close in memory); the cell order stays lexicographic (``from_device_mesh`` orders cells)."""
    device = torch.device(device)
    x = _line(n, grading, length, device)
    ni = i1 - i0
    X, Y, Z = torch.meshgrid(x[i0:i1 + 1], x, x, indexing="ij")
    points = torch.stack([X, Y, Z], dim=-1).reshape(-1, 3)
    ar = lambda m: torch.arange(m, dtype=torch.int32, device=device)  # noqa: E731
    I, J, K = torch.meshgrid(ar(ni), ar(n), ar(n), indexing="ij")
    I, J, K = I.reshape(-1), J.reshape(-1), K.reshape(-1)
    base = (I * (n + 1) + J) * (n + 1) + K                                  # local id of corner (0, 0, 0)
    off = torch.tensor([(a * (n + 1) + b) * (n + 1) + c for a in (0, 1) for b in (0, 1) for c in (0, 1)],
                       dtype=torch.int32, device=device)
    even = ((I + i0 + J + K) % 2) == 0                                      # parity of the GLOBAL hex index
    t_even = off[torch.as_tensor(_EVEN_POS, device=device)]                 # (5, 4) offsets from the hex's base corner
    t_odd = off[torch.as_tensor(_ODD_POS, device=device)]
    cells = torch.where(even[:, None, None], t_even[None], t_odd[None]) + base[:, None, None]
    cells = cells.reshape(-1, 4).contiguous()
    vertex_gid = torch.arange(points.shape[0], dtype=torch.int64, device=device) + i0 * (n + 1) * (n + 1)
    cell_gid = torch.arange(cells.shape[0], dtype=torch.int64, device=device) + 5 * i0 * n * n
    if morton_vertices:
        codes = morton_codes_device(points, torch.zeros(3, dtype=torch.float64, device=device), float(length))
        vperm = torch.argsort(codes, stable=True)
        inv = torch.empty_like(vperm, dtype=torch.int32)
        inv[vperm] = torch.arange(vperm.numel(), dtype=torch.int32, device=device)
        cells = inv[cells.long()].contiguous()
        points, vertex_gid = points[vperm], vertex_gid[vperm]
    return points.contiguous(), cells, vertex_gid, cell_gid


def hash_uniform_device(ids: torch.Tensor, seed: int = 0) -> torch.Tensor:
    """``hash_uniform`` (splitmix64 -> U[0, 1)) in int64 tensor arithmetic: products and sums wrap exactly like the
    numpy uint64 version, logical right shifts are arithmetic shifts with the sign extension masked off."""
    def i64(v):  # two's-complement view of a uint64 constant
        return v - (1 << 64) if v >= (1 << 63) else v

    def lsr(z, k):
        return (z >> k) & ((1 << (64 - k)) - 1)

    z = ids.to(torch.int64) + i64((0x9E3779B97F4A7C15 * (seed + 1)) & ((1 << 64) - 1))
    z = (z ^ lsr(z, 30)) * i64(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * i64(0x94D049BB133111EB)
    z = z ^ lsr(z, 31)
    return lsr(z, 11).to(torch.float64) * (1.0 / (1 << 53))


@dataclass
class DeviceMesh:
    """A (part of a) tet mesh resident on a device, with the global ids that make per-cell / per-vertex fields
    functions of the GLOBAL mesh (every partition of the same cube then evaluates the same model)."""

    points: torch.Tensor        # (V, 3) float64
    cells: torch.Tensor         # (T, 4) int32, local vertex ids
    vertex_gid: torch.Tensor    # (V,) int64
    cell_gid: torch.Tensor      # (T,) int64
    n_global_points: int = 0
    n_global_cells: int = 0

    @property
    def n_points(self) -> int:
        return int(self.points.shape[0])

    @property
    def n_cells(self) -> int:
        return int(self.cells.shape[0])


def cube_tet_mesh_device(n: int, device, *, grading: float = 1.0, length: float = 1.0,
                         morton_vertices: bool = True) -> DeviceMesh:
    pts, cells, vg, cg = cube_tet_slab_device(n, 0, n, device, grading=grading, length=length,
                                              morton_vertices=morton_vertices)
    return DeviceMesh(pts, cells, vg, cg, (n + 1) ** 3, 5 * n ** 3)
