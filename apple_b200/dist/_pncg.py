from __future__ import annotations

import ctypes
import math

import torch

from apple_b200 import _lib, config
from apple_b200.optim._pncg import ConvergenceCriteria, LineSearch, _classify

from ._halo import HaloExchange, PeerHaloExchange
from ._partition import Shard


class ShardedPNCG:
    """PNCG on a sharded mesh: the native phases of ``apple_b200/csrc/pncg.cu`` with the exchanges the path really
    needs between them (one halo sum per trial pass, three tiny all-reduces per iteration).  Same recurrences as the
    single-GPU fused path; vector updates are replicated on ghost vertices (their inputs are bit-identical on all
    sharers), reductions count owned entries.

    ``transport="peer"`` (default on CUDA): the exchanges are peer-memory kernels attached to the native workspace
    (``apl_pncg_set_exchange``), so ``iterate`` is ``apl_pncg_iterate`` -- the whole sharded iteration, line search
    included, runs on the device (optionally as a CUDA graph, ``use_graph`` as in the single-GPU path) and the host
    reads the scalars once per call.  ``transport="nccl"``: the phases are driven from the host with
    ``torch.distributed`` collectives between them (one tiny D2H per line-search trial); kept for backends without
    peer memory."""

    def __init__(self, potentials, ext_forces, shard: Shard, free_mask_local: torch.Tensor, u0_local: torch.Tensor,
                 *, criteria: ConvergenceCriteria | None = None, line_search: LineSearch | None = None, group=None,
                 scatter: int | None = None, transport: str | None = None, use_graph: int = 0,
                 preconditioner: str = "jacobi", psd: bool = False):
        self.shard = shard
        self.device, self.dtype = u0_local.device, u0_local.dtype
        if transport is None:
            transport = "peer" if (self.device.type == "cuda" and shard.world > 1) else "nccl"
        if shard.world == 1 and self.device.type == "cuda":
            transport = "local"        # nothing to exchange: the single-GPU device path
        self.transport = transport
        self.halo = PeerHaloExchange(shard, self.device, group) if transport == "peer" else HaloExchange(shard, self.device, group)
        self.criteria = criteria or ConvergenceCriteria()
        self.line_search = line_search or LineSearch()
        n = shard.n_local
        new = lambda: torch.zeros((n, 4), dtype=self.dtype, device=self.device)  # noqa: E731
        self.x = new()
        self.x[:, :3] = u0_local
        self.p, self.g, self.d = [new(), new()], [new(), new()], [new(), new()]
        owned = torch.as_tensor(shard.owned, device=self.device)
        mask = torch.zeros((n, 4), dtype=torch.uint8, device=self.device)
        mask[:, :3] = free_mask_local.to(torch.uint8) * (1 + 2 * owned.to(torch.uint8))[:, None]
        self.mask = mask
        self.scal = torch.zeros(_lib.PNCG_NSCAL, dtype=torch.float64, device=self.device)
        self._host = torch.zeros(_lib.PNCG_NSCAL, dtype=torch.float64).pin_memory()
        L = _lib.lib()
        handle = ctypes.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(L.apl_pncg_create(
            _lib.dtype_code(self.dtype), n, dev_index, _lib.dev_ptr(self.x), _lib.dev_ptr(self.p[0]),
            _lib.dev_ptr(self.p[1]), _lib.dev_ptr(self.g[0]), _lib.dev_ptr(self.g[1]), _lib.dev_ptr(self.d[0]),
            _lib.dev_ptr(self.d[1]), _lib.dev_ptr(self.mask), _lib.dev_ptr(self.scal), ctypes.byref(handle)))
        self._handle = handle
        self._keep = list(potentials) + list(ext_forces)
        for pot in list(potentials) + list(ext_forces):
            if pot.dtype != self.dtype:
                raise TypeError(f"potential {pot.name} is {pot.dtype}, state is {self.dtype}")
        for pot in potentials:
            _lib.check(L.apl_pncg_add_fem(handle, pot._handle))
        for ef in ext_forces:
            # Loads are applied on every rank that lists the vertex and then halo-SUMMED: a load on a shared vertex must
            # be listed by exactly one rank (its owner), or it is counted once per sharer.
            if ef.indices.numel() and not bool(owned[ef.indices.long()].all()):
                raise ValueError(f"sharded external force {ef.name}: list every loaded vertex on its OWNING rank only "
                                 "(shard.owned); a load on a ghost copy would be counted once per sharer")
            _lib.check(L.apl_pncg_add_ext_force(handle, ef.indices.shape[0], _lib.dev_ptr(ef.materials.force),
                                                _lib.dev_ptr(ef.indices)))
        c, ls = self.criteria, self.line_search
        sc = scatter if scatter is not None else config.scatter
        _lib.check(L.apl_pncg_set_params(handle, float(c.max_steps), float(c.target_relative_gradient_norm),
                                         float(c.absolute_gradient_norm), float(c.max_failed_line_searches),
                                         float(ls.overstep), 1.0, float(ls.armijo), int(ls.max_steps), int(sc),
                                         int(use_graph) if transport in ("peer", "local") else 0))
        if preconditioner not in ("jacobi", "block"):
            raise ValueError("preconditioner must be 'jacobi' or 'block'")
        self.o = [new(), new()] if preconditioner == "block" else [None, None]
        if preconditioner == "block" or psd:
            if transport not in ("peer", "local"):
                raise NotImplementedError("block Jacobi / PSD projection: peer-memory or single-GPU transport only")
            _lib.check(L.apl_pncg_set_block_jacobi(handle, _lib.dev_ptr(self.o[0]), _lib.dev_ptr(self.o[1]), int(bool(psd))))
        if transport in ("peer", "local"):
            if transport == "peer":
                _lib.check(L.apl_pncg_set_exchange(handle, self.halo._xchg))
            self._phase(_lib.PHASE_INIT)           # f, g, diag at x, completed over the ranks on the device
        else:
            self._phase(_lib.PHASE_INIT)
            self.halo.sum_(self.g[0], self.d[0])
            self.halo.all_reduce_(self.scal[_lib.S_F:_lib.S_F + 1])
        self.n_steps = 0

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.lib().apl_pncg_destroy(h)
            except Exception:
                pass
            self._handle = None

    def _phase(self, phase: int, j: int = 0) -> None:
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().apl_pncg_phase(self._handle, phase, j, _lib.stream_ptr(self.device)))

    def _read(self):
        self._host.copy_(self.scal, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._host

    def iterate(self, n_iters: int) -> float:
        """Runs up to ``n_iters`` iterations; returns the DONE code (0 = still running)."""
        if self.transport in ("peer", "local"):
            # everything on the device; iterations after DONE are no-ops on every rank
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().apl_pncg_iterate(self._handle, int(n_iters), _lib.stream_ptr(self.device)))
            s = self._read()
            self.n_steps = int(s[_lib.S_K])
            return float(s[_lib.S_DONE])
        S = _lib
        J = self.line_search.max_steps
        L = _lib.lib()
        for _ in range(n_iters):
            cur = L.apl_pncg_current(self._handle)
            trial_g, trial_d = self.g[1 - cur], self.d[1 - cur]
            self._phase(S.PHASE_REDUCE)
            self.halo.all_reduce_(self.scal[S.S_SUMS:S.S_SUMS + 11])
            self._phase(S.PHASE_FINALIZE)
            self._phase(S.PHASE_DIRECTION)
            self._phase(S.PHASE_PASS_B)
            self.halo.all_reduce_(self.scal[S.S_GP:S.S_PHP + 1])      # g.p and p.Hp are adjacent
            self._phase(S.PHASE_ALPHA)
            done = 0.0
            for j in range(J + 1):
                self._phase(S.PHASE_TRIAL, j)
                self.halo.sum_(trial_g, trial_d)
                self.halo.all_reduce_(self.scal[S.S_FT_J + j:S.S_FT_J + j + 1])
                self._phase(S.PHASE_LS, j)
                s = self._read()                                     # one tiny D2H per trial
                done = float(s[S.S_DONE])
                if done != 0.0 or float(s[S.S_ACC_J + j + 1]) != 0.0:
                    # accepted or gave up: later trials would be no-ops; carry the state to slot J+1
                    if j < J:
                        self.scal[S.S_ACC_J + j + 2:S.S_ACC_J + J + 2] = self.scal[S.S_ACC_J + j + 1]
                        self.scal[S.S_ALPHA_J + j + 2:S.S_ALPHA_J + J + 2] = self.scal[S.S_ALPHA_J + j + 1]
                    break
            self._phase(S.PHASE_COMMIT)
            _lib.check(L.apl_pncg_flip(self._handle))
            if done != 0.0:
                return done
            self.n_steps += 1
        return 0.0

    def solve(self, check_every: int = 8):
        done = 0.0
        while done == 0.0:
            done = self.iterate(check_every)
            s = self._read()
            if float(s[_lib.S_K]) >= self.criteria.max_steps and done == 0.0:
                # one more REDUCE/FINALIZE would flag it; classify here
                done = 2.0
        s = self._read()
        g0 = math.sqrt(max(float(s[_lib.S_GNORM2_FIRST]), 0.0))
        g = math.sqrt(max(float(s[_lib.S_GNORM2]), 0.0))
        rel = g / g0 if g0 > 0 else 0.0
        return {"result": _classify(done, rel, self.criteria), "n_steps": int(s[_lib.S_K]),
                "n_accepted": int(s[_lib.S_N_ACCEPTED]), "fun": float(s[_lib.S_F]), "relative_grad_norm": rel}

    @property
    def u_local(self) -> torch.Tensor:
        return self.x[:, :3].contiguous()
