from __future__ import annotations

import torch

from apple_b200 import _lib

from ._halo import HaloExchange
from ._partition import Shard


class ShardedOperators:
    """The five operators (and the fused forms) of a model whose mesh is sharded across ranks.

    ``local_model`` is any object with ``eval(ops, u, p, fun=, quad=, grad=, diag=, prod=)`` over the
    LOCAL vertex numbering of ``shard`` (a ``WarpModel`` of CUDA potentials built on ``shard.mesh``;
    the CPU tests substitute the oracle).  Inputs and outputs are local nodal fields whose shared rows
    are kept consistent across ranks; scalars are global."""

    def __init__(self, local_model, shard: Shard, device, dtype, group=None):
        self.model = local_model
        self.shard = shard
        self.halo = HaloExchange(shard, device, group)
        self.device, self.dtype = torch.device(device), dtype
        self.n_local = shard.n_local

    def eval(self, ops: int, u, p=None):
        """Returns a dict with the requested results: fun / quad (0-d tensors, global), grad / diag /
        prod ((n_local, 3), halo-summed)."""
        n = self.n_local
        new = lambda: torch.zeros((n, 3), dtype=self.dtype, device=self.device)  # noqa: E731
        out = {}
        if ops & _lib.OP_FUN:
            out["fun"] = torch.zeros(1, dtype=self.dtype, device=self.device)
        if ops & _lib.OP_HESS_QUAD:
            out["quad"] = torch.zeros(1, dtype=self.dtype, device=self.device)
        if ops & _lib.OP_GRAD:
            out["grad"] = new()
        if ops & _lib.OP_HESS_DIAG:
            out["diag"] = new()
        if ops & _lib.OP_HESS_PROD:
            out["prod"] = new()
        self.model.eval(ops, u, p, **out)
        fields = [out[k] for k in ("grad", "diag", "prod") if k in out]
        self.halo.sum_(*fields)
        scalars = [out[k] for k in ("fun", "quad") if k in out]
        if len(scalars) == 1:
            self.halo.all_reduce_(scalars[0])
            for k in ("fun", "quad"):
                if k in out:
                    out[k] = out[k][0]
        elif scalars:
            s = torch.cat(scalars)
            self.halo.all_reduce_(s)
            for i, k in enumerate(k for k in ("fun", "quad") if k in out):
                out[k] = s[i]
        return out

    def fun_grad_hess_prod(self, u, p):
        r = self.eval(_lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_PROD, u, p)
        return r["fun"], r["grad"], r["prod"]
