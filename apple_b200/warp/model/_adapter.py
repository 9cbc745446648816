from __future__ import annotations

import torch

from apple_b200 import _lib

from ._model import WarpModel


def zeros_block(n_points: int, n_fields: int, n_scalars: int, dtype: torch.dtype, device):
    """One allocation holding ``n_fields`` nodal fields ``(n_points, 3)`` (each starting at a multiple of 16
    bytes, as the vector REDs of the kernels require) followed by ``n_scalars`` one-element tensors.  Zeroing
    the outputs of a fused evaluation is then ONE memset (``buf.zero_()``) instead of one per output."""
    stride = (3 * n_points + 3) // 4 * 4
    buf = torch.zeros(n_fields * stride + max(n_scalars, 1), dtype=dtype, device=device)
    fields = [buf[k * stride:k * stride + 3 * n_points].view(n_points, 3) for k in range(n_fields)]
    scalars = [buf[n_fields * stride + k:n_fields * stride + k + 1] for k in range(n_scalars)]
    return buf, fields, scalars


class WarpModelAdapter:
    """Pure-function view of a ``WarpModel``, ``warp/model/_adapter.py:16-140``.

    The reference wraps five JAX-FFI callables; here the same five functions allocate their output
    on ``u``'s device and call straight into the CUDA kernels on the current torch stream.  dtype is
    taken from ``u`` like the reference's dtype-generic factories (``_adapter.py:46-47``)."""

    def __init__(self, model: WarpModel, n_points: int):
        self.__wrapped__ = model
        self.n_points = int(n_points)

    def _scalar(self, u: torch.Tensor) -> torch.Tensor:
        return torch.empty(1, dtype=u.dtype, device=u.device)

    def _field(self, u: torch.Tensor) -> torch.Tensor:
        return torch.empty((self.n_points, 3), dtype=u.dtype, device=u.device)

    def fun(self, u: torch.Tensor) -> torch.Tensor:
        u = u.contiguous()
        output = self._scalar(u)
        self.__wrapped__.fun(u, output)
        return output[0]

    def grad(self, u: torch.Tensor) -> torch.Tensor:
        u = u.contiguous()
        output = self._field(u)
        self.__wrapped__.grad(u, output)
        return output

    def hess_diag(self, u: torch.Tensor) -> torch.Tensor:
        u = u.contiguous()
        output = self._field(u)
        self.__wrapped__.hess_diag(u, output)
        return output

    def hess_prod(self, u: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
        u, p = u.contiguous(), p.contiguous()
        output = self._field(u)
        self.__wrapped__.hess_prod(u, p, output)
        return output

    def hess_quad(self, u: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
        u, p = u.contiguous(), p.contiguous()
        output = self._scalar(u)
        self.__wrapped__.hess_quad(u, p, output)
        return output[0]

    # ---- fused forms (one pass over the elements per potential) ----
    def fun_grad_hess_prod(self, u: torch.Tensor, p: torch.Tensor, *, scatter=None):
        """(energy, gradient, Hessian-vector product): the fused evaluation of the headline metric."""
        u, p = u.contiguous(), p.contiguous()
        _, (grad, prod), (fun,) = zeros_block(self.n_points, 2, 1, u.dtype, u.device)   # one memset for all outputs
        self.__wrapped__.eval(
            _lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_PROD, u, p, fun=fun, grad=grad, prod=prod, scatter=scatter,
            zero=False,
        )
        return fun[0], grad, prod

    def fun_grad_hess_prod_host(self, u: torch.Tensor, p: torch.Tensor, out=None, *, scatter=None):
        """Host-buffer form of :meth:`fun_grad_hess_prod`: ``u`` and ``p`` are HOST tensors ``(n_points, 3)``
        (pinned memory makes the copies asynchronous), the results land in the host tensors
        ``out = (fun (1,), grad (n_points, 3), prod (n_points, 3))`` (allocated pinned and reused when
        ``None``).  This is what a caller whose state lives in host memory pays per evaluation, so the
        two PCIe directions are overlapped with the element passes instead of running
        copy-in -> kernel -> copy-out back to back:

            copy stream in :  u ---------> p --------->
            compute stream :              [fun+grad(u)]   [hess_prod(u, p)]
            copy stream out:                           grad, fun ------>     prod ------>

        Energy and gradient need only ``u`` and start while ``p`` is still in flight; the gradient
        travels back while the Hessian-vector pass runs.  Two element passes cost more GPU time than
        the fused one, but they hide behind the interconnect, which is the bottleneck of this call.
        Everything is ordered after prior work on the current stream, and the current stream waits for
        the last copy: synchronising it (or any later work on it) sees the host results."""
        model = self.__wrapped__
        st = self._host_state(u.dtype)
        dev = st["device"]
        n = self.n_points
        if out is None:
            out = st["out"]
        fun_h, grad_h, prod_h = out
        cur = torch.cuda.current_stream(dev)
        s_in, s_out = st["s_in"], st["s_out"]
        s_in.wait_stream(cur)     # earlier kernels on the current stream may still read the staging buffers
        with torch.cuda.stream(s_in):
            st["u"].copy_(u.reshape(n, 3), non_blocking=True)
            ev_u = s_in.record_event()
            st["p"].copy_(p.reshape(n, 3), non_blocking=True)
            ev_p = s_in.record_event()
        with torch.cuda.device(dev):
            cur.wait_event(ev_u)
            st["block_a"].zero_()
            model.eval(_lib.OP_FUN | _lib.OP_GRAD, st["u"], None, fun=st["fun"], grad=st["grad"], scatter=scatter,
                       zero=False)
            s_out.wait_stream(cur)
            with torch.cuda.stream(s_out):
                grad_h.copy_(st["grad"], non_blocking=True)
                fun_h.copy_(st["fun"], non_blocking=True)
            cur.wait_event(ev_p)
            model.eval(_lib.OP_HESS_PROD, st["u"], st["p"], prod=st["prod"], scatter=scatter)
            s_out.wait_stream(cur)
            with torch.cuda.stream(s_out):
                prod_h.copy_(st["prod"], non_blocking=True)
            cur.wait_stream(s_out)
        return fun_h, grad_h, prod_h

    def _host_state(self, dtype: torch.dtype) -> dict:
        cache = self.__dict__.setdefault("_host_cache", {})
        if dtype not in cache:
            pots = list(self.__wrapped__.potentials.values())
            dev = torch.device(getattr(pots[0], "device", "cuda")) if pots else torch.device("cuda")
            if dev.type != "cuda":
                raise _lib.NativeError("fun_grad_hess_prod_host needs a model on a CUDA device (there is no CPU path)")
            n = self.n_points
            field = lambda: torch.empty((n, 3), dtype=dtype, device=dev)  # noqa: E731
            block_a, (grad,), (fun,) = zeros_block(n, 1, 1, dtype, dev)   # outputs of the first pass: one memset
            cache[dtype] = {
                "device": dev, "u": field(), "p": field(), "grad": grad, "prod": field(), "fun": fun, "block_a": block_a,
                "s_in": torch.cuda.Stream(dev), "s_out": torch.cuda.Stream(dev),
                "out": (torch.empty(1, dtype=dtype).pin_memory(), torch.empty((n, 3), dtype=dtype).pin_memory(),
                        torch.empty((n, 3), dtype=dtype).pin_memory()),
            }
        return cache[dtype]

    def fun_grad_hess_diag(self, u: torch.Tensor, *, scatter=None):
        """(energy, gradient, Hessian diagonal): PNCG's pass A."""
        u = u.contiguous()
        _, (grad, diag), (fun,) = zeros_block(self.n_points, 2, 1, u.dtype, u.device)
        self.__wrapped__.eval(
            _lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_DIAG, u, None, fun=fun, grad=grad, diag=diag, scatter=scatter,
            zero=False,
        )
        return fun[0], grad, diag
