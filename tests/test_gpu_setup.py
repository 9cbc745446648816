"""GPU: device-side setup (apl_fem_create_from_mesh -- Morton keys, radix sort, rest shape and static planes computed
on the GPU) against the host path (Region.compute_grad + apl_fem_create) and against the oracle."""

import ctypes

import numpy as np
import pytest
import torch

from helpers import KINDS, cuda_potential, make_case, oracle_potential, rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1.0e-5, torch.float64: 1.0e-10}


def _device_potential(kind, mesh, dtype, **kw):
    from apple_b200.warp.fem import Arap, StableNeoHookean, StableNeoHookeanMuscle

    cells = torch.as_tensor(np.ascontiguousarray(mesh.cells, dtype=np.int32), device="cuda")
    points = torch.as_tensor(mesh.points, dtype=torch.float64, device="cuda")
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype, device="cuda")  # noqa: E731
    mats = {"mu": t(mesh.cell_data["mu"])}
    if kind != "arap":
        mats["lambda_"] = t(mesh.cell_data["lambda"])
    if kind == "muscle":
        mats["activation"] = t(mesh.cell_data["activation"])
    cls = {"snh": StableNeoHookean, "arap": Arap, "muscle": StableNeoHookeanMuscle}[kind]
    return cls.from_device_mesh(cells, points, fraction=t(mesh.cell_data["Fraction"]), dtype=dtype, **mats, **kw)


def _host_order(pot):
    from apple_b200 import _lib

    L = _lib.lib()
    info = (ctypes.c_int64 * 10)()
    L.apl_fem_info(pot._handle, info)
    order = np.zeros(info[9], np.int64)
    tiles = np.zeros((info[2], 6), np.int32)
    L.apl_fem_host_tables(pot._handle, _lib.host_ptr(tiles), _lib.host_ptr(order), None, None, None, None, None)
    return order, tiles


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("kind", KINDS)
def test_device_setup_matches_host_setup_and_oracle(native_lib, kind, dtype):
    from apple_b200 import _lib

    mesh, u, p = make_case(n=9, seed=5)          # 3645 tets: 15 tiles
    V = mesh.n_points
    host = cuda_potential(kind, mesh, dtype)
    dev = _device_potential(kind, mesh, dtype)
    # same Morton keys, same stable order, hence the same tiles
    oh, th = _host_order(host)
    od, td = _host_order(dev)
    assert np.array_equal(oh, od) and np.array_equal(th, td)
    ora = oracle_potential(kind, mesh)
    ud = torch.as_tensor(u, dtype=dtype, device="cuda")
    pd = torch.as_tensor(p, dtype=dtype, device="cuda")
    ops = _lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_DIAG | _lib.OP_HESS_PROD | _lib.OP_HESS_QUAD
    res = {}
    for name, pot in (("host", host), ("dev", dev)):
        o = {"fun": torch.zeros(1, dtype=dtype, device="cuda"), "quad": torch.zeros(1, dtype=dtype, device="cuda"),
             "grad": torch.zeros((V, 3), dtype=dtype, device="cuda"), "diag": torch.zeros((V, 3), dtype=dtype, device="cuda"),
             "prod": torch.zeros((V, 3), dtype=dtype, device="cuda")}
        pot.eval(ops, ud, pd, **o)
        torch.cuda.synchronize()
        res[name] = {k: v.cpu().numpy() for k, v in o.items()}
    ref = {"fun": np.zeros(1), "quad": np.zeros(1), "grad": np.zeros((V, 3)), "diag": np.zeros((V, 3)), "prod": np.zeros((V, 3))}
    ora.fun(u, ref["fun"]); ora.hess_quad(u, p, ref["quad"]); ora.grad(u, ref["grad"])
    ora.hess_diag(u, ref["diag"]); ora.hess_prod(u, p, ref["prod"])
    for k in ref:
        assert rel_err(res["dev"][k], ref[k]) < TOL[dtype], k
        assert rel_err(res["dev"][k], res["host"][k]) < TOL[dtype], k
    # mixed derivative products use the device-resident packed order of this path
    if kind != "arap":
        mh, md = host.mixed_derivative_prod(ud, pd), dev.mixed_derivative_prod(ud, pd)
        for k in mh:
            assert rel_err(md[k].cpu(), mh[k].cpu()) < 10 * TOL[dtype], k


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_device_setup_fused_pair_and_given_order(native_lib, dtype):
    from apple_b200 import _lib
    from apple_b200.warp.fem import FusedSnhArap, StableNeoHookean

    mesh, u, p = make_case(n=8, seed=6)
    V, T = mesh.n_points, mesh.n_cells
    m2 = mesh.copy()
    m2.cell_data["Fraction"] = 1.0 - 0.5 * mesh.cell_data["Fraction"]
    m2.cell_data["mu"] = mesh.cell_data["mu"][::-1].copy()
    from oracle import fem as ofem

    ora = ofem.Model([oracle_potential("snh", mesh), oracle_potential("arap", m2)], V)
    cells = torch.as_tensor(np.ascontiguousarray(mesh.cells, dtype=np.int32), device="cuda")
    points = torch.as_tensor(mesh.points, dtype=torch.float64, device="cuda")
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype, device="cuda")  # noqa: E731
    fused = FusedSnhArap.from_device_mesh(cells, points, mu=t(mesh.cell_data["mu"]), lambda_=t(mesh.cell_data["lambda"]),
                                          fraction=t(mesh.cell_data["Fraction"]), mu_arap=t(m2.cell_data["mu"]),
                                          fraction_arap=t(m2.cell_data["Fraction"]), dtype=dtype)
    ud = torch.as_tensor(u, dtype=dtype, device="cuda"); pd = torch.as_tensor(p, dtype=dtype, device="cuda")
    fun = torch.zeros(1, dtype=dtype, device="cuda")
    grad = torch.zeros((V, 3), dtype=dtype, device="cuda"); prod = torch.zeros((V, 3), dtype=dtype, device="cuda")
    fused.eval(_lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_PROD, ud, pd, fun=fun, grad=grad, prod=prod)
    torch.cuda.synchronize()
    assert rel_err(fun.cpu(), ora.fun(u)) < TOL[dtype]
    assert rel_err(grad.cpu(), ora.grad(u)) < TOL[dtype]
    assert rel_err(prod.cpu(), ora.hess_prod(u, p)) < TOL[dtype]
    # morton=False keeps the caller's cell order
    snh = StableNeoHookean.from_device_mesh(cells, points, mu=t(mesh.cell_data["mu"]), lambda_=t(mesh.cell_data["lambda"]),
                                            dtype=dtype, morton=False)
    order, _ = _host_order(snh)
    assert np.array_equal(order, np.arange(T))


def test_device_setup_rejects_bad_meshes(native_lib):
    from apple_b200 import _lib
    from apple_b200.warp.fem import Arap

    mesh, _, _ = make_case(n=3, seed=0)
    cells = torch.as_tensor(np.ascontiguousarray(mesh.cells, dtype=np.int32), device="cuda")
    points = torch.as_tensor(mesh.points, dtype=torch.float64, device="cuda")
    mu = torch.ones(mesh.n_cells, device="cuda")
    bad = cells.clone(); bad[5, 2] = mesh.n_points          # index out of range
    with pytest.raises(_lib.NativeError, match="outside"):
        Arap.from_device_mesh(bad, points, mu=mu, dtype=torch.float32)
    flat = cells.clone(); flat[7, 3] = flat[7, 0]           # repeated vertex: zero rest volume
    with pytest.raises(_lib.NativeError, match="degenerate"):
        Arap.from_device_mesh(flat, points, mu=mu, dtype=torch.float32)
    with pytest.raises(_lib.NativeError):
        Arap.from_device_mesh(cells.cpu(), points.cpu(), mu=mu, dtype=torch.float32)
