from __future__ import annotations

import torch

from apple_b200 import _lib

from apple_b200.warp.model._adapter import zeros_block

from ._halo import HaloExchange, PeerHaloExchange
from ._partition import Shard


class ShardedOperators:
    """The five operators (and the fused forms) of a model whose mesh is sharded across ranks.

    ``local_model`` is any object with ``eval(ops, u, p, fun=, quad=, grad=, diag=, prod=)`` over the
    LOCAL vertex numbering of ``shard`` (a ``WarpModel`` of CUDA potentials built on ``shard.mesh``;
    the CPU tests substitute the oracle).  Inputs and outputs are local nodal fields whose shared rows
    are kept consistent across ranks; scalars are global."""

    def __init__(self, local_model, shard: Shard, device, dtype, group=None, overlap: bool = True,
                 transport: str | None = None, peer_overlap: bool = False):
        """``transport``: ``"peer"`` = halo sums and the scalar reduction through peer memory over NVLink in two small
        kernels per evaluation (``PeerHaloExchange``; CUDA, one node), ``"nccl"`` = pack / all-to-all / unpack /
        all-reduce through ``torch.distributed`` (any backend; with ``overlap`` the exchange runs while the interior
        tiles are evaluated).  Default: peer memory on CUDA devices, ``torch.distributed`` otherwise."""
        self.model = local_model
        self.shard = shard
        self.device, self.dtype = torch.device(device), dtype
        if transport is None:
            transport = "peer" if (self.device.type == "cuda" and shard.world > 1) else "nccl"
        self.transport = transport
        self.halo = PeerHaloExchange(shard, device, group) if transport == "peer" else HaloExchange(shard, device, group)
        # peer transport: the exchange is two small kernels behind the element pass.  The boundary-first form below
        # (peer_overlap=True) is OPT-IN: measured at 2 GPUs / 64 M tets it is slower (1.367 vs 1.339 ms per step, run
        # r2z3): the small boundary launch and the second, scalar-only exchange cost more than the hidden wait.
        peer_overlap = transport == "peer" and peer_overlap
        if transport == "peer" and not peer_overlap:
            overlap = False
        self.n_local = shard.n_local
        # Split evaluation: element tiles that touch a shared vertex first, then the halo exchange of their
        # results on a side stream WHILE the interior tiles (which touch no shared vertex) are evaluated.
        self._split = hasattr(local_model, "mark_boundary")
        self.overlap = False
        self.n_boundary_tiles = 0
        # peer transport: boundary tiles first, PUSH their shared rows, interior tiles, PULL -- the peers' rows and flags
        # arrive while the interior pass runs, so the pull does not wait (and rank skew is absorbed); the partial scalars,
        # complete only after the interior pass, follow in a second, tiny exchange
        self.peer_overlap = False
        if overlap and self._split and shard.world > 1 and self.halo.total > 0:
            import numpy as np

            flags = np.zeros(shard.n_local, np.uint8)
            for idx in shard.neighbors.values():
                flags[idx] = 1
            self.n_boundary_tiles = int(local_model.mark_boundary(flags))
            self.overlap = not peer_overlap
            self.peer_overlap = peer_overlap
        # high priority: the pack / exchange / unpack kernels become ready together with the interior pass and
        # must get SM resources first, otherwise the persistent interior kernel would simply run ahead of them
        self._side = torch.cuda.Stream(self.device, priority=-1) if self.device.type == "cuda" else None

    def eval(self, ops: int, u, p=None):
        """Returns a dict with the requested results: fun / quad (0-d tensors, global), grad / diag /
        prod ((n_local, 3), halo-summed)."""
        n = self.n_local
        names = [k for k, bit in (("grad", _lib.OP_GRAD), ("diag", _lib.OP_HESS_DIAG), ("prod", _lib.OP_HESS_PROD)) if ops & bit]
        snames = [k for k, bit in (("fun", _lib.OP_FUN), ("quad", _lib.OP_HESS_QUAD)) if ops & bit]
        # one allocation + one memset for all outputs; every field starts at a multiple of 16 bytes
        buf, fields, scalars = zeros_block(n, len(names), len(snames), self.dtype, self.device)
        out = dict(zip(names, fields))
        out.update(zip(snames, scalars))
        scal = buf[buf.numel() - max(len(snames), 1):]     # the scalars are contiguous at the end of the block
        kw = {"zero": False} if self._split else {}    # the outputs above are already zero
        if self.overlap and fields:
            self.model.eval(ops, u, p, part=_lib.PART_BOUNDARY, **out, **kw)
            if self._side is not None:
                cur = torch.cuda.current_stream(self.device)
                self._side.wait_stream(cur)
                with torch.cuda.stream(self._side):
                    self.halo.sum_(*fields)            # pack, all-to-all, unpack: shared rows only
                self.model.eval(ops, u, p, part=_lib.PART_INTERIOR, **out, **kw)
                cur.wait_stream(self._side)
            else:                                      # CPU (gloo tests): same order, no overlap
                self.halo.sum_(*fields)
                self.model.eval(ops, u, p, part=_lib.PART_INTERIOR, **out, **kw)
        elif self.peer_overlap and fields:
            self.model.eval(ops, u, p, part=_lib.PART_BOUNDARY, **out, **kw)
            self.halo.push(fields)
            self.model.eval(ops, u, p, part=_lib.PART_INTERIOR, **out, **kw)
            self.halo.pull(fields)
            if snames:
                self.halo.all_reduce_(scal)
                for i, k in enumerate(snames):
                    out[k] = scal[i]
            return out
        elif self.transport == "peer":
            # one element pass, then ONE push + ONE pull kernel carry the shared rows and the partial scalars
            self.model.eval(ops, u, p, **out, **kw)
            self.halo.sum_(*fields, scal=scal if snames else None)
            for i, k in enumerate(snames):
                out[k] = scal[i]
            return out
        else:
            self.model.eval(ops, u, p, **out, **kw)
            self.halo.sum_(*fields)
        if snames:
            self.halo.all_reduce_(scal)
            for i, k in enumerate(snames):
                out[k] = scal[i]
        return out

    def fun_grad_hess_prod(self, u, p):
        r = self.eval(_lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_PROD, u, p)
        return r["fun"], r["grad"], r["prod"]
