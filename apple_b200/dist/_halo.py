from __future__ import annotations

import torch
import torch.distributed as dist

from ._partition import Shard


class HaloExchange:
    """Sums partial nodal values over all ranks that share a vertex.

    One ``all_to_all_single`` per call (NCCL over NVLink on GPUs, gloo in the CPU tests) moves only
    the shared rows.  Every rank then adds the partials of a vertex in ascending RANK order (its own
    partial in its own position), so all replicas of a vertex end up bit-identical."""

    def __init__(self, shard: Shard, device, group=None):
        self.shard = shard
        self.group = group
        self.world, self.rank = shard.world, shard.rank
        self.device = torch.device(device)
        self.idx = {s: torch.as_tensor(ix, dtype=torch.int64, device=self.device) for s, ix in shard.neighbors.items()}
        self.counts = [int(shard.neighbors[s].size) if s in shard.neighbors else 0 for s in range(self.world)]
        self.total = sum(self.counts)
        offs = [0]
        for c in self.counts:
            offs.append(offs[-1] + c)
        self.offs = offs
        if self.total:
            self.send_index = torch.cat([self.idx[s] for s in range(self.world) if s in self.idx])
            shared = torch.unique(self.send_index)
        else:
            self.send_index = torch.zeros(0, dtype=torch.int64, device=self.device)
            shared = self.send_index
        self.shared = shared  # local ids of all vertices shared with anybody

    def sum_(self, *fields: torch.Tensor) -> None:
        """In place: every field (n_local, k) gets, on shared rows, the sum over all sharers."""
        if self.world == 1 or not fields:
            return
        width = [f.shape[1] for f in fields]
        cat = fields[0] if len(fields) == 1 else torch.cat(fields, dim=1)
        send = cat.index_select(0, self.send_index).contiguous()
        recv = torch.empty_like(send)
        splits = self.counts
        dist.all_to_all_single(recv, send, output_split_sizes=splits, input_split_sizes=splits, group=self.group)
        own = cat.index_select(0, self.shared)
        acc = torch.zeros_like(cat)
        for s in range(self.world):  # ascending rank order on every rank
            if s == self.rank:
                acc.index_add_(0, self.shared, own)
            elif self.counts[s]:
                acc.index_add_(0, self.idx[s], recv[self.offs[s]:self.offs[s + 1]])
        col = 0
        for f, w in zip(fields, width):
            f.index_copy_(0, self.shared, acc.index_select(0, self.shared)[:, col:col + w])
            col += w

    def all_reduce_(self, t: torch.Tensor) -> None:
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
