from __future__ import annotations

from collections.abc import Mapping

import torch

from ._potential import WarpPotential


class WarpModel:
    """Sum of potentials, ``warp/model/_model.py:9-36``: zero the output, then let every potential
    accumulate into it."""

    def __init__(self, potentials: Mapping[str, WarpPotential]):
        self.potentials = dict(potentials)

    def fun(self, u: torch.Tensor, output: torch.Tensor) -> None:
        output.zero_()
        for potential in self.potentials.values():
            potential.fun(u, output)

    def grad(self, u: torch.Tensor, output: torch.Tensor) -> None:
        output.zero_()
        for potential in self.potentials.values():
            potential.grad(u, output)

    def hess_diag(self, u: torch.Tensor, output: torch.Tensor) -> None:
        output.zero_()
        for potential in self.potentials.values():
            potential.hess_diag(u, output)

    def hess_prod(self, u: torch.Tensor, p: torch.Tensor, output: torch.Tensor) -> None:
        output.zero_()
        for potential in self.potentials.values():
            potential.hess_prod(u, p, output)

    def hess_quad(self, u: torch.Tensor, p: torch.Tensor, output: torch.Tensor) -> None:
        output.zero_()
        for potential in self.potentials.values():
            potential.hess_quad(u, p, output)

    def hess_prod_psd(self, u: torch.Tensor, p: torch.Tensor, output: torch.Tensor) -> None:
        """Opt-in superset: the sum of the potentials' PSD-projected Hessian-vector products (``APL_OP_PSD``)."""
        output.zero_()
        for potential in self.potentials.values():
            if hasattr(potential, "hess_prod_psd"):
                potential.hess_prod_psd(u, p, output)

    def mixed_derivative_prod(self, u: torch.Tensor, p: torch.Tensor) -> dict:
        """``{potential name: {material name: d/dq [grad E . p] per cell}}`` over the potentials that have materials."""
        return {name: pot.mixed_derivative_prod(u, p) for name, pot in self.potentials.items()
                if hasattr(pot, "mixed_derivative_prod")}

    def mark_boundary(self, vertex_flags) -> int:
        """Forwards to every potential that has element tiles; returns the total number of boundary tiles."""
        total = 0
        for potential in self.potentials.values():
            if hasattr(potential, "mark_boundary"):
                total += int(potential.mark_boundary(vertex_flags))
        return total

    def eval(self, ops: int, u, p=None, *, fun=None, quad=None, grad=None, diag=None, prod=None, scatter=None,
             part: int = 0, zero: bool = True) -> None:
        """Fused form: every requested operator of every potential in one pass per potential.  ``part`` / ``zero``
        serve split evaluations (boundary tiles first, interior tiles while the halo exchange runs): the second
        call passes ``zero=False`` and accumulates."""
        if zero:
            for out in (fun, quad, grad, diag, prod):
                if out is not None:
                    out.zero_()
        for potential in self.potentials.values():
            if part:
                potential.eval(ops, u, p, fun=fun, quad=quad, grad=grad, diag=diag, prod=prod, scatter=scatter, part=part)
            else:
                potential.eval(ops, u, p, fun=fun, quad=quad, grad=grad, diag=diag, prod=prod, scatter=scatter)
