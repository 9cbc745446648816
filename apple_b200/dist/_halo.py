from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from apple_b200 import _lib

from ._partition import Shard


class HaloExchange:
    """Sums partial nodal values over all ranks that share a vertex.

    Per call: one pack kernel, one ``all_to_all_single`` of the shared rows only (NCCL over NVLink on
    GPUs, gloo in the CPU tests), one unpack kernel.  Every rank adds the partials of a vertex in
    ascending RANK order (its own partial in its own position), so all replicas of a vertex end up
    bit-identical.  On CUDA tensors pack/unpack are the ``apl_halo_*`` kernels of the native library;
    the torch index-op implementation below is kept for the CPU (gloo) tests of the sharding logic."""

    def __init__(self, shard: Shard, device, group=None):
        self.shard = shard
        self.group = group
        self.world, self.rank = shard.world, shard.rank
        self.device = torch.device(device)
        nb = shard.neighbors
        self.idx = {s: torch.as_tensor(ix, dtype=torch.int64, device=self.device) for s, ix in nb.items()}
        self.counts = [int(nb[s].size) if s in nb else 0 for s in range(self.world)]
        self.total = sum(self.counts)
        offs = [0]
        for c in self.counts:
            offs.append(offs[-1] + c)
        self.offs = offs
        send_index = np.concatenate([nb[s] for s in range(self.world) if s in nb]) if self.total else np.zeros(0, np.int64)
        shared = np.unique(send_index)
        # CSR over shared vertices: contributions in ascending rank order, -1 marks this rank's own partial
        verts = [shared]
        ranks = [np.full(shared.size, self.rank, np.int64)]
        srcs = [np.full(shared.size, -1, np.int64)]
        for s in range(self.world):
            if s != self.rank and s in nb:
                verts.append(nb[s])
                ranks.append(np.full(nb[s].size, s, np.int64))
                srcs.append(offs[s] + np.arange(nb[s].size, dtype=np.int64))
        verts, ranks, srcs = np.concatenate(verts), np.concatenate(ranks), np.concatenate(srcs)
        order = np.lexsort((ranks, verts))              # by vertex, then by rank
        src = srcs[order]
        counts = np.bincount(np.searchsorted(shared, verts), minlength=shared.size)
        row_ptr = np.zeros(shared.size + 1, np.int32)
        row_ptr[1:] = np.cumsum(counts)
        t = lambda a, dt: torch.as_tensor(a, dtype=dt, device=self.device)  # noqa: E731
        self.send_index = t(send_index, torch.int64)
        self.shared = t(shared, torch.int64)
        self.row_ptr = t(row_ptr, torch.int32)
        self.src = t(src, torch.int64)
        self._buf = {}

    def _buffers(self, nf, dtype):
        key = (nf, dtype)
        if key not in self._buf:
            self._buf[key] = (torch.empty((self.total, nf * 3), dtype=dtype, device=self.device),
                              torch.empty((self.total, nf * 3), dtype=dtype, device=self.device))
        return self._buf[key]

    def sum_(self, *fields: torch.Tensor) -> None:
        """In place: the first three columns of every field (n_local, 3|4) get, on shared rows, the
        sum over all sharers."""
        if self.world == 1 or not fields or self.total == 0:
            return
        if len(fields) > 3:
            self.sum_(*fields[:3])
            self.sum_(*fields[3:])
            return
        nf, dtype, ld = len(fields), fields[0].dtype, int(fields[0].shape[1])
        send, recv = self._buffers(nf, dtype)
        if fields[0].is_cuda:
            L = _lib.lib()
            f = [_lib.dev_ptr(x) for x in fields] + [None] * (3 - nf)
            with torch.cuda.device(self.device):
                st = _lib.stream_ptr(self.device)
                _lib.check(L.apl_halo_pack(_lib.dtype_code(dtype), self.total, _lib.dev_ptr(self.send_index), nf, f[0], f[1],
                                           f[2], ld, _lib.dev_ptr(send), st))
                dist.all_to_all_single(recv, send, output_split_sizes=self.counts, input_split_sizes=self.counts,
                                       group=self.group)
                _lib.check(L.apl_halo_unpack(_lib.dtype_code(dtype), self.shared.numel(), _lib.dev_ptr(self.shared),
                                             _lib.dev_ptr(self.row_ptr), _lib.dev_ptr(self.src), nf, f[0], f[1], f[2], ld,
                                             _lib.dev_ptr(recv), st))
            return
        # CPU tensors (gloo tests of the sharding logic): the same pack / CSR-ordered unpack as the kernels,
        # written with torch index ops
        for k, x in enumerate(fields):
            send[:, 3 * k:3 * k + 3] = x[:, :3].index_select(0, self.send_index)
        dist.all_to_all_single(recv, send, output_split_sizes=self.counts, input_split_sizes=self.counts,
                               group=self.group)
        n_shared = self.shared.numel()
        counts = (self.row_ptr[1:] - self.row_ptr[:-1]).to(torch.int64)
        entry_row = torch.repeat_interleave(torch.arange(n_shared), counts)
        pos = torch.arange(self.src.numel()) - self.row_ptr[:-1].to(torch.int64)[entry_row]
        own_entry = self.src < 0
        for k, x in enumerate(fields):
            contrib = torch.empty((self.src.numel(), 3), dtype=dtype)
            contrib[own_entry] = x[:, :3].index_select(0, self.shared[entry_row[own_entry]])
            contrib[~own_entry] = recv[self.src[~own_entry], 3 * k:3 * k + 3]
            acc = torch.zeros((n_shared, 3), dtype=dtype)
            for j in range(int(counts.max()) if n_shared else 0):   # entry j of every row: ascending rank order
                sel = pos == j
                acc.index_add_(0, entry_row[sel], contrib[sel])
            x[:, :3].index_copy_(0, self.shared, acc)

    def all_reduce_(self, t: torch.Tensor) -> None:
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
