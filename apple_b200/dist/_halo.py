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


class PeerHaloExchange(HaloExchange):
    """``HaloExchange`` whose data path is peer memory over NVLink (``apl_xchg_*``, ``csrc/xchg.cu``): the shared rows
    of up to three fields AND the partial scalars of a step travel in ONE push kernel (plain stores into the sharers'
    receive buffers) followed by ONE pull kernel (device-side wait on epoch flags, rank-ordered sums), instead of
    pack + NCCL all-to-all + unpack + NCCL all-reduce.  Same plan, same summation order, bit-identical results.
    ``torch.distributed`` (any backend) is used once, at construction, to exchange the plan sizes and the CUDA IPC
    handles.  One process per GPU, all ranks on one node."""

    def __init__(self, shard: Shard, device, group=None):
        super().__init__(shard, device, group)
        if self.device.type != "cuda":
            raise _lib.NativeError("PeerHaloExchange needs CUDA devices (peer memory over NVLink)")
        import ctypes

        L = _lib.lib()
        world, rank = self.world, self.rank
        # counts[r][s] = rows rank r shares with rank s: rank s's receive buffer holds the segments of ranks 0..world-1
        # in rank order, so this rank's rows start at sum(counts[s][:rank]) there
        all_counts = [None] * world
        dist.all_gather_object(all_counts, list(self.counts), group=group)
        peers, rows = [], []
        for s in range(world):
            if s == rank or self.counts[s] == 0:
                continue
            if all_counts[s][rank] != self.counts[s]:
                raise RuntimeError(f"halo plan mismatch between ranks {rank} and {s}")
            base = sum(all_counts[s][:rank])
            peers.append(np.full(self.counts[s], s, np.int32))
            rows.append(base + np.arange(self.counts[s], dtype=np.int64))
        t = lambda a, dt: torch.as_tensor(a, dtype=dt, device=self.device)  # noqa: E731
        self.send_peer = t(np.concatenate(peers) if peers else np.zeros(0, np.int32), torch.int32)
        self.send_row = t(np.concatenate(rows) if rows else np.zeros(0, np.int64), torch.int64)
        handle = ctypes.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(self.device):
            # the SAME capacity on every rank: a push addresses the peer's second receive buffer with the local buffer
            # size (csrc/xchg.cu), so regions of different sizes (end ranks of a slab partition share one plane, middle
            # ranks two) would be written out of bounds
            max_rows = max(max(sum(c) for c in all_counts), 1)
            _lib.check(L.apl_xchg_create(world, rank, dev_index, max_rows, ctypes.byref(handle)))
            self._xchg = handle
            mine = ctypes.create_string_buffer(64)
            _lib.check(L.apl_xchg_ipc_handle(handle, mine))
            blobs = [None] * world
            dist.all_gather_object(blobs, bytes(mine.raw), group=group)
            _lib.check(L.apl_xchg_connect(handle, ctypes.create_string_buffer(b"".join(blobs), 64 * world)))
            _lib.check(L.apl_xchg_set_plan(handle, self.total, _lib.dev_ptr(self.send_index), _lib.dev_ptr(self.send_peer),
                                           _lib.dev_ptr(self.send_row), self.shared.numel(), _lib.dev_ptr(self.shared),
                                           _lib.dev_ptr(self.row_ptr), _lib.dev_ptr(self.src)))
        dist.barrier(group=group)       # every region is mapped everywhere before the first push

    def __del__(self):
        h = getattr(self, "_xchg", None)
        if h is not None and h.value:
            try:
                _lib.lib().apl_xchg_destroy(h)
            except Exception:  # interpreter shutdown
                pass
            self._xchg = None

    def _call(self, fn, fields, scal):
        nf = len(fields)
        if nf > 3:
            raise ValueError("at most three fields per exchange")
        dtype = fields[0].dtype if nf else scal.dtype
        ld = int(fields[0].shape[1]) if nf else 3
        f = [_lib.dev_ptr(x) for x in fields] + [None] * (3 - nf)
        n_scal = 0 if scal is None else int(scal.numel())
        with torch.cuda.device(self.device):
            _lib.check(fn(self._xchg, _lib.dtype_code(dtype), nf, f[0], f[1], f[2], ld, _lib.dev_ptr(scal), n_scal,
                          _lib.stream_ptr(self.device)))

    def push(self, fields, scal=None) -> None:
        self._call(_lib.lib().apl_xchg_push, fields, scal)

    def pull(self, fields, scal=None) -> None:
        self._call(_lib.lib().apl_xchg_pull, fields, scal)

    def sum_(self, *fields: torch.Tensor, scal: torch.Tensor | None = None) -> None:
        """Halo sum of ``fields`` (in place) and, in the same exchange, the global sums of the entries of ``scal``."""
        if self.world == 1:
            return
        for k in range(0, max(len(fields), 1), 3):
            part = fields[k:k + 3]
            s = scal if k == 0 else None
            if not part and s is None:
                continue
            self.push(part, s)
            self.pull(part, s)

    def all_reduce_(self, t: torch.Tensor) -> None:
        if self.world > 1:
            self.push((), t)
            self.pull((), t)
