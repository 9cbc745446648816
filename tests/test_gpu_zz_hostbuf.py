"""GPU: the host-buffer entry point of the adapter (PCIe copies overlapped with two element passes on side
streams).  Kept in its own late-running file: it exercises stream choreography that the operator parity
tests do not depend on."""

import numpy as np
import pytest
import torch

from helpers import cuda_potential, make_case, oracle_potential, rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1.0e-5, torch.float64: 1.0e-10}


@pytest.fixture(scope="module")
def case():
    return make_case(n=7, seed=3)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_host_buffer_entry_point_matches_device_path(native_lib, case, dtype):
    """WarpModelAdapter.fun_grad_hess_prod_host (host tensors in, host tensors out, copies overlapped
    with two element passes on side streams) == the oracle and the fused device-tensor call; repeated
    calls with changing inputs reuse the staging buffers without races."""
    from apple_b200.mesh import lumped_vertex_volume
    from apple_b200.warp.fem import fuse_potentials
    from apple_b200.warp.model import WarpModel, WarpModelAdapter
    from apple_b200.warp.potential import ExternalForce
    from oracle import fem as ofem

    mesh, u, p = case
    V = mesh.n_points
    idx = np.arange(0, V, 3)
    force = np.zeros((idx.size, 3)); force[:, 2] = -9.8 * 1000.0 * lumped_vertex_volume(mesh)[idx]
    pots = fuse_potentials({k: cuda_potential(k, mesh, dtype, name=k) for k in ("snh", "arap")})
    pots["force"] = ExternalForce(idx, force, dtype=dtype, name="force")
    adapter = WarpModelAdapter(WarpModel(pots), n_points=V)
    omodel = ofem.Model([oracle_potential("snh", mesh), oracle_potential("arap", mesh), ofem.ExternalForce(force, idx)], V)
    tol = 2 * TOL[dtype]
    uh = torch.as_tensor(u, dtype=dtype).pin_memory(); ph = torch.as_tensor(p, dtype=dtype).pin_memory()
    for rep, scale in enumerate((1.0, 2.0, 0.5)):
        uh.copy_(torch.as_tensor(scale * u, dtype=dtype)); ph.copy_(torch.as_tensor(p / scale, dtype=dtype))
        f, g, h = adapter.fun_grad_hess_prod_host(uh, ph)
        torch.cuda.current_stream().synchronize()
        assert not g.is_cuda and g.shape == (V, 3)
        assert rel_err(f, omodel.fun(scale * u)) < tol, rep
        assert rel_err(g, omodel.grad(scale * u)) < tol, rep
        assert rel_err(h, omodel.hess_prod(scale * u, p / scale)) < tol, rep
        fd, gd, hd = adapter.fun_grad_hess_prod(uh.cuda(), ph.cuda())
        assert rel_err(g, gd.cpu()) < tol and rel_err(h, hd.cpu()) < tol and rel_err(f, fd.cpu()) < tol
    # caller-provided (pageable) outputs work too
    out = (torch.empty(1, dtype=dtype), torch.empty((V, 3), dtype=dtype), torch.empty((V, 3), dtype=dtype))
    f2, g2, h2 = adapter.fun_grad_hess_prod_host(uh, ph, out=out)
    torch.cuda.synchronize()
    assert g2 is out[1] and rel_err(g2, g) < tol and rel_err(h2, h) < tol
