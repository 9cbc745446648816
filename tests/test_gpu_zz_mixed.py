"""GPU: apl_fem_mixed_derivative_prod (d/dq [grad E . p] per cell; the reference's historical
model.mixed_derivative_prod) against central differences of the oracle.  The closed forms are also validated on the
CPU (tests/test_native_cpu.py)."""

import contextlib
import copy

import numpy as np
import pytest
import torch

from helpers import KINDS, cuda_potential, make_case, oracle_potential

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-6), (torch.float32, 2e-4)], ids=["f64", "f32"])
@pytest.mark.parametrize("kind", KINDS)
def test_mixed_derivative_prod_matches_finite_differences(native_lib, kind, dtype, tol):
    from apple_b200.warp.model import WarpModel

    mesh, u, p = make_case(n=5, seed=2, amp=0.15)
    ora = oracle_potential(kind, mesh)
    pot = cuda_potential(kind, mesh, dtype, name=kind)
    ud = torch.as_tensor(u, dtype=dtype, device="cuda"); pd = torch.as_tensor(p, dtype=dtype, device="cuda")
    got = WarpModel({kind: pot}).mixed_derivative_prod(ud, pd)[kind]
    torch.cuda.synchronize()

    def phi(o):
        return np.einsum("cai,cai->c", o.elem_grad(u), p[mesh.cells])

    def fd(name, comp=None, eps=1e-6):
        lo, hi = copy.deepcopy(ora), copy.deepcopy(ora)
        for o, sign in ((hi, 1.0), (lo, -1.0)):
            m = dict(ora.materials)
            x = np.array(ora.materials[name], dtype=float, copy=True)
            if comp is None:
                x += sign * eps
            else:
                x[:, comp] += sign * eps
            m[name] = x
            o.materials = m
        return (phi(hi) - phi(lo)) / (2 * eps)

    for name, val in got.items():
        val = val.cpu().numpy().astype(np.float64)
        if name == "activation":
            for k in range(6):
                ref = fd(name, k)
                assert np.abs(val[:, k] - ref).max() <= tol * np.abs(ref).max(), (name, k)
        else:
            ref = fd(name)
            assert np.abs(val - ref).max() <= tol * np.abs(ref).max(), name
