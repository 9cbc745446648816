"""CPU: the plain-C restatement (oracle/c, used as the timed CPU baseline) equals the numpy oracle."""

import numpy as np
import pytest

from helpers import KINDS, make_case, oracle_potential, rel_err


@pytest.mark.parametrize("kind", KINDS)
def test_c_oracle_matches_numpy_oracle(kind):
    from oracle import cbind

    mesh, u, p = make_case(n=5, seed=6, amp=0.2)
    V = mesh.n_points
    ref = oracle_potential(kind, mesh)
    cp = cbind.CPotential(kind, ref.cells, ref.dhdX, ref.dV, ref.materials["mu"], ref.materials.get("lambda_"),
                          ref.materials.get("activation"))
    for name, args in (("fun", (u,)), ("hess_quad", (u, p))):
        a, b = np.zeros(1), np.zeros(1)
        getattr(cp, name)(*args, a); getattr(ref, name)(*args, b)
        assert rel_err(a, b) < 1e-11, name
    for name, args in (("grad", (u,)), ("hess_diag", (u,)), ("hess_prod", (u, p))):
        a, b = np.zeros((V, 3)), np.zeros((V, 3))
        getattr(cp, name)(*args, a); getattr(ref, name)(*args, b)
        assert rel_err(a, b) < 1e-11, name
    assert cbind.num_threads() >= 1
