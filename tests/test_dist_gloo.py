"""CPU, world_size = 2 over gloo: the sharding logic (partition, halo plans, ordered halo sums, scalar
all-reduce) with the oracle standing in for the per-rank element operators.  The assembled results of
the two ranks must equal the single-rank oracle on the whole mesh."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import make_case, oracle_potential


class _OracleLocalModel:
    """Adapts oracle potentials (numpy) to the ``eval`` interface ShardedOperators drives, including the
    split evaluation (``mark_boundary`` / ``part``): here the boundary part is the set of CELLS touching a
    flagged vertex (the CUDA potentials split by tiles; either way the interior part touches no flagged
    vertex, which is the property the overlap relies on)."""

    def __init__(self, pots, n_points):
        from oracle import fem as ofem

        self.pots, self.n_points = pots, n_points
        self.m = {0: ofem.Model(pots, n_points)}

    def mark_boundary(self, flags):
        import copy

        from oracle import fem as ofem

        flags = np.asarray(flags).astype(bool)
        parts = {1: [], 2: []}
        n_boundary = 0
        for pot in self.pots:
            touch = flags[pot.cells].any(axis=1)
            n_boundary += int(touch.sum())
            for part, sel in ((1, touch), (2, ~touch)):
                sub = copy.copy(pot)
                sub.cells, sub.dhdX, sub.dV = pot.cells[sel], pot.dhdX[sel], pot.dV[sel]
                sub.materials = {k: np.asarray(v)[sel] for k, v in pot.materials.items()}
                parts[part].append(sub)
        for part in (1, 2):       # a thin slab may have no interior cells at all
            nonempty = [q for q in parts[part] if len(q.cells)]
            self.m[part] = ofem.Model(nonempty, self.n_points) if nonempty else None
        return n_boundary

    def eval(self, ops, u, p=None, *, fun=None, quad=None, grad=None, diag=None, prod=None, scatter=None, part=0,
             zero=True):
        m = self.m[part]
        if zero:
            for o in (fun, quad, grad, diag, prod):
                if o is not None:
                    o.zero_()
        if m is None:
            return
        zero = False
        un = u.numpy()
        pn = None if p is None else p.numpy()
        if zero:
            for o in (fun, quad, grad, diag, prod):
                if o is not None:
                    o.zero_()
        if fun is not None:
            fun += float(m.fun(un))
        if quad is not None:
            quad += float(m.hess_quad(un, pn))
        if grad is not None:
            grad += torch.from_numpy(m.grad(un))
        if diag is not None:
            diag += torch.from_numpy(m.hess_diag(un))
        if prod is not None:
            prod += torch.from_numpy(m.hess_prod(un, pn))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, overlap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from apple_b200 import _lib
        from apple_b200.dist import ShardedOperators, partition_mesh

        mesh, u, p = make_case(n=5, seed=4)
        shard = partition_mesh(mesh, world, rank)
        pots = [oracle_potential(k, shard.mesh) for k in ("snh", "arap")]
        ops = ShardedOperators(_OracleLocalModel(pots, shard.n_local), shard, "cpu", torch.float64, overlap=overlap)
        assert ops.overlap == overlap and (ops.n_boundary_tiles > 0) == overlap
        ul = torch.from_numpy(u[shard.l2g]).contiguous()
        pl = torch.from_numpy(p[shard.l2g]).contiguous()
        r = ops.eval(_lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_DIAG | _lib.OP_HESS_PROD | _lib.OP_HESS_QUAD, ul, pl)
        out[rank] = {
            "l2g": shard.l2g, "owned": shard.owned, "fun": float(r["fun"]), "quad": float(r["quad"]),
            "grad": r["grad"].numpy(), "diag": r["diag"].numpy(), "prod": r["prod"].numpy(),
            "n_neighbors": len(shard.neighbors), "cells": shard.cell_range,
        }
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("overlap", [False, True], ids=["serial", "split"])
@pytest.mark.parametrize("world", [2, 4])
def test_sharding_matches_single_rank(world, overlap):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out, overlap), nprocs=world, join=True)
    from oracle import fem as ofem

    mesh, u, p = make_case(n=5, seed=4)
    V = mesh.n_points
    ref = ofem.Model([oracle_potential(k, mesh) for k in ("snh", "arap")], V)
    g, d, h = ref.grad(u), ref.hess_diag(u), ref.hess_prod(u, p)
    e, q = ref.fun(u), ref.hess_quad(u, p)
    owners = np.zeros(V, int)
    touched = np.zeros(V, int)
    for r in range(world - 1):
        assert out[r]["cells"][1] == out[r + 1]["cells"][0]    # contiguous, disjoint tet chunks
    for r in range(world):
        o = out[r]
        l2g = o["l2g"]
        touched[l2g] += 1
        owners[l2g[o["owned"]]] += 1
        assert o["n_neighbors"] >= 1
        assert abs(o["fun"] - e) <= 1e-12 * abs(e) and abs(o["quad"] - q) <= 1e-12 * abs(q)
        for name, full in (("grad", g), ("diag", d), ("prod", h)):
            # every local copy (owned AND ghost) holds the global sum
            assert np.abs(o[name] - full[l2g]).max() <= 1e-12 * np.abs(full).max(), name
    assert (touched >= 1).all() and (owners == 1).all()        # every vertex has exactly one owner
    if world > 2:
        assert (touched >= 3).any()                            # some vertices have three or more sharers
    # replicas of shared vertices are bit-identical across ranks (sums taken in ascending rank order)
    for ra in range(world):
        for rb in range(ra + 1, world):
            a, b = out[ra], out[rb]
            common, ia, ib = np.intersect1d(a["l2g"], b["l2g"], return_indices=True)
            for name in ("grad", "diag", "prod"):
                assert np.array_equal(a[name][ia], b[name][ib])


def test_partition_is_deterministic_and_covers_four_ranks():
    from apple_b200.dist import partition_mesh

    mesh, _, _ = make_case(n=6, seed=1)
    shards = [partition_mesh(mesh, 4, r) for r in range(4)]
    assert sum(s.mesh.n_cells for s in shards) == mesh.n_cells
    owned = np.zeros(mesh.n_points, int)
    for s in shards:
        owned[s.l2g[s.owned]] += 1
        assert np.array_equal(mesh.cells[s.cell_range[0]:s.cell_range[1]], s.l2g[s.mesh.cells])
        for t, idx in s.neighbors.items():
            other = shards[t]
            # both sides list the shared vertices in the same (global id) order
            assert np.array_equal(s.l2g[idx], other.l2g[other.neighbors[s.rank]])
    assert (owned == 1).all()


def _slab_worker(rank, world, port, out, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from apple_b200 import _lib
        from apple_b200.dist import ShardedOperators, slab_shard
        from apple_b200.mesh import hash_uniform
        from oracle import fem as ofem
        from oracle import region

        shard = slab_shard(n, world, rank)                       # this rank's slab only: no global mesh
        m = shard.mesh
        cg, vg = m.cell_data["gid"], m.point_data["gid"]
        mu = 1.0 + 2.0 * hash_uniform(cg, 1); la = 1.0 + 8.0 * hash_uniform(cg, 2)      # functions of the GLOBAL ids
        u = 0.01 * (np.stack([hash_uniform(3 * vg + k, 3) for k in range(3)], 1) - 0.5)
        p = np.stack([hash_uniform(3 * vg + k, 4) for k in range(3)], 1) - 0.5
        dhdX, dV = region.compute_grad(m.points, m.cells)
        pots = [ofem.StableNeoHookean(m.cells, dhdX, dV, mu=mu, lambda_=la), ofem.Arap(m.cells, dhdX, dV, mu=mu)]
        ops = ShardedOperators(_OracleLocalModel(pots, shard.n_local), shard, "cpu", torch.float64, overlap=True)
        r = ops.eval(_lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_PROD, torch.from_numpy(u), torch.from_numpy(p))
        out[rank] = {"gid": vg, "owned": shard.owned, "fun": float(r["fun"]), "grad": r["grad"].numpy(),
                     "prod": r["prod"].numpy(), "n_cells": m.n_cells, "cells": shard.cell_range}
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 3])
def test_slab_shards_generated_per_rank_match_the_whole_cube(world):
    """slab_shard: every rank generates ONLY its hex layers of the cube (no global mesh), fields are functions of the
    global ids; the sharded result equals the single-rank oracle on the whole cube."""
    from apple_b200.mesh import cube_tet_mesh, hash_uniform
    from oracle import fem as ofem
    from oracle import region

    n = 5
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_slab_worker, args=(world, _free_port(), out, n), nprocs=world, join=True)
    full = cube_tet_mesh(n, morton=False)                        # global lexicographic mesh == the ids of the slabs
    T, V = full.n_cells, full.n_points
    cg, vg = np.arange(T), np.arange(V)
    mu = 1.0 + 2.0 * hash_uniform(cg, 1); la = 1.0 + 8.0 * hash_uniform(cg, 2)
    u = 0.01 * (np.stack([hash_uniform(3 * vg + k, 3) for k in range(3)], 1) - 0.5)
    p = np.stack([hash_uniform(3 * vg + k, 4) for k in range(3)], 1) - 0.5
    dhdX, dV = region.compute_grad(full.points, full.cells)
    ref = ofem.Model([ofem.StableNeoHookean(full.cells, dhdX, dV, mu=mu, lambda_=la), ofem.Arap(full.cells, dhdX, dV, mu=mu)], V)
    e, g, h = ref.fun(u), ref.grad(u), ref.hess_prod(u, p)
    owners = np.zeros(V, int)
    assert sum(out[r]["n_cells"] for r in range(world)) == T
    for r in range(world):
        o = out[r]
        owners[o["gid"][o["owned"]]] += 1
        assert abs(o["fun"] - e) <= 1e-12 * abs(e)
        assert np.abs(o["grad"] - g[o["gid"]]).max() <= 1e-12 * np.abs(g).max()
        assert np.abs(o["prod"] - h[o["gid"]]).max() <= 1e-12 * np.abs(h).max()
    assert (owners == 1).all()
