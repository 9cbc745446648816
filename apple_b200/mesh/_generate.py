"""Synthetic meshes for tests and benchmarks (all seeded / deterministic).

The reference builds its grids with VTK (``exp/2025/12/31/inverse-toy/src/10-gen-grid.py:35-42``:
``to_tetrahedra(tetra_per_cell=5)`` plus a winding fix).  Here the conforming 5-tet split of a
structured hex grid is generated directly: hexes of even parity use one diagonal pattern, odd parity
the mirrored one, and every tet is oriented so that its rest volume is positive
(``jax/fem/region/_region.py:98-99`` only warns on ``dV <= 0``).
"""

from __future__ import annotations

import numpy as np

from ._mesh import TetMesh

# corner index = 4*dx + 2*dy + dz
_EVEN = np.array([[0, 1, 2, 4], [3, 1, 7, 2], [5, 1, 4, 7], [6, 2, 7, 4], [1, 2, 4, 7]])
_ODD = np.array([[1, 0, 3, 5], [2, 0, 6, 3], [4, 0, 5, 6], [7, 3, 6, 5], [0, 3, 5, 6]])


def _part1by2(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64) & np.uint64(0x1FFFFF)
    x = (x | (x << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    x = (x | (x << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    return x


def morton_codes(xyz: np.ndarray) -> np.ndarray:
    """63-bit Morton code of points (normalised to their bounding cube)."""
    lo = xyz.min(axis=0)
    ext = float((xyz.max(axis=0) - lo).max())
    scale = ((1 << 21) - 1) / ext if ext > 0 else 0.0
    q = ((xyz - lo) * scale).astype(np.uint64)
    return _part1by2(q[:, 0]) | (_part1by2(q[:, 1]) << np.uint64(1)) | (_part1by2(q[:, 2]) << np.uint64(2))


def morton_reorder(mesh: TetMesh, *, first_touch: bool = True) -> TetMesh:
    """Orders the tets along a Morton curve and renumbers the vertices to match (data arrays are
    permuted with them).  ``first_touch=True`` numbers vertices in the order the Morton-sorted tets
    first reference them, so that the vertices a run of consecutive tets touches for the first time
    are contiguous in memory (fewer cache lines per gather); otherwise vertices follow their own
    Morton code."""
    cperm = np.argsort(morton_codes(mesh.points[mesh.cells].mean(axis=1)), kind="stable")
    cells = mesh.cells[cperm]
    if first_touch:
        flat = cells.reshape(-1)
        _, first = np.unique(flat, return_index=True)       # first occurrence of every referenced vertex
        touched = flat[np.sort(first)]
        rest = np.setdiff1d(np.arange(mesh.n_points), touched, assume_unique=False)
        vperm = np.concatenate([touched, rest])
    else:
        vperm = np.argsort(morton_codes(mesh.points), kind="stable")
    inv = np.empty(mesh.n_points, dtype=np.int64)
    inv[vperm] = np.arange(vperm.size)
    points = mesh.points[vperm]
    cells = inv[cells].astype(np.int32)
    point_data = {k: np.asarray(v)[vperm] for k, v in mesh.point_data.items()}
    cell_data = {k: np.asarray(v)[cperm] for k, v in mesh.cell_data.items()}
    return TetMesh(points, cells, point_data, cell_data)


def cube_tet_mesh(n, *, grading: float = 1.0, length: float = 1.0, morton: bool = True) -> TetMesh:
    """Unit-cube (``length``) grid of ``n``^3 hexes (or ``(nx, ny, nz)``) split into 5 tets each.

    ``grading`` > 1 spaces the grid lines geometrically (ratio per layer) so that ``dhdX`` and ``dV``
    vary from tet to tet.  ``morton=True`` numbers vertices and tets along a Morton curve."""
    nx, ny, nz = (n, n, n) if np.isscalar(n) else n

    def line(m):
        if grading == 1.0:
            return np.linspace(0.0, length, m + 1)
        w = grading ** np.arange(m)
        x = np.concatenate([[0.0], np.cumsum(w)])
        return x / x[-1] * length

    X, Y, Z = np.meshgrid(line(nx), line(ny), line(nz), indexing="ij")
    points = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)

    def vid(i, j, k):
        return (i * (ny + 1) + j) * (nz + 1) + k

    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    corners = np.stack([vid(I + a, J + b, K + c) for a in (0, 1) for b in (0, 1) for c in (0, 1)], axis=1)
    even = ((I + J + K) % 2) == 0
    cells = np.where(even[:, None, None], corners[:, _EVEN], corners[:, _ODD]).reshape(-1, 4)
    # orient: positive rest volume
    X4 = points[cells]
    vol = np.einsum("ci,ci->c", np.cross(X4[:, 1] - X4[:, 0], X4[:, 2] - X4[:, 0]), X4[:, 3] - X4[:, 0])
    neg = vol < 0
    cells[neg] = cells[neg][:, [0, 2, 1, 3]]
    mesh = TetMesh(points, cells.astype(np.int32))
    return morton_reorder(mesh) if morton else mesh


def cube_tet_slab(n: int, i0: int, i1: int, *, grading: float = 1.0, length: float = 1.0):
    """The hex layers ``i0 <= i < i1`` (first grid axis) of ``cube_tet_mesh(n, morton=False)``, built WITHOUT the rest
    of the cube: ``(points, cells, vertex_gid, cell_gid)`` with local vertex numbering, the global vertex ids
    ``(i (n+1) + j) (n+1) + k`` and the global cell ids ``5 hex + t`` of the full lexicographic mesh.  A rank of a
    sharded run generates only its slab (a 64 M-tet cube is 25 GB of numpy arrays as a whole)."""
    def line(m):
        if grading == 1.0:
            return np.linspace(0.0, length, m + 1)
        w = grading ** np.arange(m)
        x = np.concatenate([[0.0], np.cumsum(w)])
        return x / x[-1] * length

    x = line(n)
    X, Y, Z = np.meshgrid(x[i0:i1 + 1], x, x, indexing="ij")
    points = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    ni = i1 - i0

    def vid(i, j, k):          # local id of grid vertex (i0 + i, j, k)
        return (i * (n + 1) + j) * (n + 1) + k

    I, J, K = np.meshgrid(np.arange(ni), np.arange(n), np.arange(n), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    corners = np.stack([vid(I + a, J + b, K + c) for a in (0, 1) for b in (0, 1) for c in (0, 1)], axis=1)
    even = ((I + i0 + J + K) % 2) == 0                      # parity of the GLOBAL hex index
    cells = np.where(even[:, None, None], corners[:, _EVEN], corners[:, _ODD]).reshape(-1, 4)
    X4 = points[cells]
    vol = np.einsum("ci,ci->c", np.cross(X4[:, 1] - X4[:, 0], X4[:, 2] - X4[:, 0]), X4[:, 3] - X4[:, 0])
    neg = vol < 0
    cells[neg] = cells[neg][:, [0, 2, 1, 3]]
    vertex_gid = np.arange(points.shape[0], dtype=np.int64) + i0 * (n + 1) * (n + 1)
    cell_gid = np.arange(cells.shape[0], dtype=np.int64) + 5 * i0 * n * n
    return points, cells.astype(np.int32), vertex_gid, cell_gid


def hash_uniform(ids, seed: int = 0) -> np.ndarray:
    """Deterministic U[0, 1) per integer id (splitmix64): the same value for a global vertex / cell id on every
    rank and for every partition, without generating the global array."""
    with np.errstate(over="ignore"):
        z = np.asarray(ids, dtype=np.uint64) + np.full(1, 0x9E3779B97F4A7C15, np.uint64) * np.uint64(seed + 1)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def embedded_tetra_mesh() -> TetMesh:
    """The reference's known-answer mesh: 5 points, 4 tets around an interior vertex
    (``tests/forward/test_static_simulation.py:16-55``)."""
    points = np.array(
        [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [0.25, 0.25, 0.25]]
    )
    cells = np.array([4, 0, 1, 2, 4, 4, 0, 1, 4, 3, 4, 0, 4, 2, 3, 4, 4, 1, 2, 3], dtype=np.int64)
    return TetMesh(points, cells)


def lumped_vertex_volume(mesh: TetMesh) -> np.ndarray:
    """Per-vertex lumped volume (a quarter of every incident tet), for gravity-like loads."""
    X4 = mesh.points[mesh.cells]
    vol = np.einsum("ci,ci->c", np.cross(X4[:, 1] - X4[:, 0], X4[:, 2] - X4[:, 0]), X4[:, 3] - X4[:, 0]) / 6.0
    out = np.zeros(mesh.n_points)
    np.add.at(out, mesh.cells.reshape(-1), np.repeat(vol / 4.0, 4))
    return out
