from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from apple_b200.mesh import TetMesh


@dataclass
class Shard:
    """One rank's part of a partitioned mesh."""

    rank: int
    world: int
    mesh: TetMesh                    # local mesh (local vertex numbering), cell/point data carried over
    l2g: np.ndarray                  # (n_local,) local -> global vertex id (ascending)
    owned: np.ndarray                # (n_local,) bool: this rank owns the vertex (lowest rank touching it)
    cell_range: tuple[int, int]      # [lo, hi) of the global (ordered) cells held here
    # halo plan: for every other rank s that shares vertices with this rank, the LOCAL indices of the
    # shared vertices, sorted by global id (both sides list them in the same order)
    neighbors: dict[int, np.ndarray] = field(default_factory=dict)
    n_global_points: int = 0
    n_global_cells: int = 0

    @property
    def n_local(self) -> int:
        return int(self.l2g.shape[0])


def partition_mesh(mesh: TetMesh, world: int, rank: int) -> Shard:
    """Contiguous chunks of the cell order (use a Morton-ordered mesh, ``mesh.morton_reorder``).

    Deterministic and identical on every rank (each rank computes the full ownership table from the
    connectivity; at 64M tets this is a few seconds of numpy and is setup-time work)."""
    T, V = mesh.n_cells, mesh.n_points
    if not 1 <= world <= 64:
        raise ValueError("partition_mesh keeps the ranks touching a vertex in a 64-bit mask: 1 <= world <= 64")
    bounds = [r * T // world for r in range(world + 1)]
    lo, hi = bounds[rank], bounds[rank + 1]
    cells = mesh.cells[lo:hi]
    l2g = np.unique(cells)
    g2l = np.full(V, -1, dtype=np.int64)
    g2l[l2g] = np.arange(l2g.size)
    local = TetMesh(
        mesh.points[l2g],
        g2l[cells].astype(np.int32),
        point_data={k: np.asarray(v)[l2g] for k, v in mesh.point_data.items()},
        cell_data={k: np.asarray(v)[lo:hi] for k, v in mesh.cell_data.items()},
    )
    # which ranks touch each vertex: bit mask per vertex (world <= 64)
    touch = np.zeros(V, dtype=np.uint64)
    for r in range(world):
        vr = np.unique(mesh.cells[bounds[r]:bounds[r + 1]])
        touch[vr] |= np.uint64(1) << np.uint64(r)
    mine = touch[l2g]
    low_bit = mine & (~mine + np.uint64(1))          # lowest set bit = owner
    owned = low_bit == (np.uint64(1) << np.uint64(rank))
    neighbors = {}
    for s in range(world):
        if s == rank:
            continue
        shared = np.flatnonzero((mine >> np.uint64(s)) & np.uint64(1))
        if shared.size:
            neighbors[s] = shared.astype(np.int64)   # l2g ascending => sorted by global id
    return Shard(rank=rank, world=world, mesh=local, l2g=l2g, owned=owned, cell_range=(lo, hi),
                 neighbors=neighbors, n_global_points=V, n_global_cells=T)


def slab_shard(n: int, world: int, rank: int, *, grading: float = 1.0, morton: bool = True) -> Shard:
    """Rank ``rank``'s part of the ``n``^3 x 5 cube, generated WITHOUT the global mesh: the hex layers
    ``[rank n / world, (rank + 1) n / world)`` along the first grid axis (``cube_tet_slab``), locally Morton-ordered.
    The halo plan is analytic: a rank shares exactly the grid plane ``i = i1`` with the next rank and ``i = i0`` with
    the previous one; the lower rank owns the shared plane.  ``shard.mesh.point_data["gid"]`` /
    ``cell_data["gid"]`` carry the global vertex / cell ids of the full lexicographic mesh, so that per-vertex and
    per-cell fields can be defined as functions of the global id (``apple_b200.mesh.hash_uniform``) and agree across
    partitions.  Used for the strong-scaling sweep on meshes too large to build on every rank."""
    from apple_b200.mesh import cube_tet_slab, morton_reorder

    if world > n:
        raise ValueError("more ranks than hex layers")
    bounds = [r * n // world for r in range(world + 1)]
    i0, i1 = bounds[rank], bounds[rank + 1]
    points, cells, vgid, cgid = cube_tet_slab(n, i0, i1, grading=grading)
    mesh = TetMesh(points, cells, point_data={"gid": vgid}, cell_data={"gid": cgid})
    if morton:
        mesh = morton_reorder(mesh)
    vgid = np.asarray(mesh.point_data["gid"], dtype=np.int64)
    plane = (n + 1) * (n + 1)
    layer = vgid // plane                                 # grid index i of every local vertex
    neighbors = {}
    for other, shared_layer in ((rank - 1, i0), (rank + 1, i1)):
        if 0 <= other < world:
            idx = np.flatnonzero(layer == shared_layer)
            neighbors[other] = idx[np.argsort(vgid[idx], kind="stable")].astype(np.int64)   # both sides: by global id
    owned = np.ones(vgid.size, dtype=bool)
    if rank > 0:
        owned[layer == i0] = False                        # the plane shared with the previous rank is owned by it
    return Shard(rank=rank, world=world, mesh=mesh, l2g=vgid, owned=owned,
                 cell_range=(5 * i0 * n * n, 5 * i1 * n * n), neighbors=neighbors,
                 n_global_points=(n + 1) ** 3, n_global_cells=5 * n ** 3)


def slab_shard_device(n: int, world: int, rank: int, device, *, grading: float = 1.0) -> Shard:
    """``slab_shard`` with the mesh GENERATED ON THE DEVICE (``apple_b200.mesh.cube_tet_slab_device``): ``shard.mesh`` is
    a ``DeviceMesh`` (torch tensors in HBM: points, cells, global vertex / cell ids), ``l2g`` and ``owned`` are device
    tensors, and only the analytic halo plan (the one or two shared grid planes, a few 10^4 indices) visits the host.
    Local vertices are numbered along a Morton curve.  For meshes that should never exist in host memory (config 5:
    64 M tets)."""
    import torch

    from apple_b200.mesh import DeviceMesh, cube_tet_slab_device

    if world > n:
        raise ValueError("more ranks than hex layers")
    bounds = [r * n // world for r in range(world + 1)]
    i0, i1 = bounds[rank], bounds[rank + 1]
    pts, cells, vgid, cgid = cube_tet_slab_device(n, i0, i1, device, grading=grading, morton_vertices=True)
    mesh = DeviceMesh(pts, cells, vgid, cgid, (n + 1) ** 3, 5 * n ** 3)
    plane = (n + 1) * (n + 1)
    layer = torch.div(vgid, plane, rounding_mode="floor")
    neighbors = {}
    for other, shared_layer in ((rank - 1, i0), (rank + 1, i1)):
        if 0 <= other < world:
            idx = torch.nonzero(layer == shared_layer).reshape(-1)
            idx = idx[torch.argsort(vgid[idx], stable=True)]                     # both sides: by global id
            neighbors[other] = idx.cpu().numpy().astype(np.int64)
    owned = torch.ones(vgid.numel(), dtype=torch.bool, device=vgid.device)
    if rank > 0:
        owned[layer == i0] = False                                               # the previous rank owns the shared plane
    return Shard(rank=rank, world=world, mesh=mesh, l2g=vgid, owned=owned,
                 cell_range=(5 * i0 * n * n, 5 * i1 * n * n), neighbors=neighbors,
                 n_global_points=(n + 1) ** 3, n_global_cells=5 * n ** 3)
