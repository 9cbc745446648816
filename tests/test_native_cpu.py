"""CPU: the native library loads and exports its ABI; host-side tiling; the per-tet math header
compiled for the host against the oracle.  No CUDA call is made here."""

import ctypes
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from helpers import KINDS, make_case, oracle_potential

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(native_lib):
    from apple_b200 import _lib

    header = (ROOT / "include" / "apple_b200.h").read_text()
    declared = set(re.findall(r"\b(apl_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(native_lib, name), name
    assert native_lib.apl_version() >= 100


def test_no_gpu_means_loud_failure(native_lib):
    """On a machine without a GPU every product entry point must raise, never fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from apple_b200 import NativeError
    from apple_b200.mesh import embedded_tetra_mesh
    from apple_b200.warp.fem import Arap

    mesh = embedded_tetra_mesh()
    mesh.cell_data["mu"] = np.ones(4)
    with pytest.raises(NativeError):
        Arap.from_pyvista(mesh)
    assert native_lib.apl_device_count() < 0


def _host_tables(native_lib, mesh, kind=0, with_points=True):
    from apple_b200 import _lib
    from oracle import region

    dhdX, dV = region.compute_grad(mesh.points, mesh.cells)
    T, V = mesh.n_cells, mesh.n_points
    mu = np.ones(T); la = np.ones(T); act = np.zeros((T, 6))
    h = ctypes.c_void_p()
    P = _lib.host_ptr
    cells = np.ascontiguousarray(mesh.cells, dtype=np.int32)
    pts = np.ascontiguousarray(mesh.points) if with_points else None
    rc = native_lib.apl_fem_create(kind, _lib.F64, T, V, P(cells), P(dhdX), P(dV), P(mu), P(la), P(act), P(pts), -1,
                                   ctypes.byref(h))
    assert rc == 0, native_lib.apl_last_error()
    info = (ctypes.c_int64 * 10)()
    native_lib.apl_fem_info(h, info)
    nt, nv, nvo = info[2], info[3], info[8]
    tiles = np.zeros((nt, 6), np.int32); order = np.zeros(T, np.int64)
    conn = np.zeros((T, 4), np.uint8); slots = np.zeros((T, 4), np.uint16)
    tv = np.zeros(nv, np.int32); voff = np.zeros(nvo, np.uint16); vperm = np.zeros(nv, np.uint8)
    native_lib.apl_fem_host_tables(h, P(tiles), P(order), P(conn), P(slots), P(tv), P(voff), P(vperm))
    native_lib.apl_fem_destroy(h)
    return tiles, order, conn, slots, tv, voff, vperm


@pytest.mark.parametrize("with_points", [True, False])
def test_tiling_invariants(native_lib, with_points):
    mesh, _, _ = make_case(n=9, seed=0, morton=False)
    tiles, order, conn, slots, tv, voff, vperm = _host_tables(native_lib, mesh, with_points=with_points)
    T = mesh.n_cells
    assert sorted(order.tolist()) == list(range(T))          # a permutation of the cells
    if not with_points:
        assert (order == np.arange(T)).all()                 # NULL points keeps the caller's order
    assert tiles[:, 1].sum() == T and (tiles[:, 1] <= 256).all() and (tiles[:, 3] <= 256).all()
    assert (tiles[:, 1] == 256).sum() >= len(tiles) - 2       # a regular mesh fills its tiles
    assert (tiles[1:, 0] == tiles[:-1, 0] + tiles[:-1, 1]).all()
    assert (tiles[:, 0] % 4 == 0).all() and (tiles[:, 2] % 16 == 0).all() and (tiles[:, 4] % 8 == 0).all()
    for t, (ts, n, vs, nv, vo, nslots) in enumerate(tiles):
        gl = tv[vs:vs + nv]
        assert len(set(gl.tolist())) == nv                   # distinct global ids, indexed by local id
        assert np.array_equal(gl[conn[ts:ts + n]], mesh.cells[order[ts:ts + n]])   # connectivity round trip
        raw = voff[vo: vo + nv + 1].astype(int)
        start, pad = raw & 0x0fff, raw[:-1] >> 12             # reduce order; bits 12..15 = pad slots after the range
        cnt = np.diff(start) - pad                            # valence of the t-th vertex in reduce order
        perm = vperm[vs:vs + nv].astype(int)
        assert sorted(perm.tolist()) == list(range(nv))      # a permutation of the local ids ...
        assert cnt.sum() == 4 * n and (cnt > 0).all() and (pad <= 3).all()
        assert start[-1] == nslots <= 4 * 256 + 192           # fits the kernels' slot buffer
        gmax = np.array([cnt[g:g + 16].max() for g in range(0, nv, 16)])
        gmin = np.array([cnt[g:g + 16].min() for g in range(0, nv, 16)])
        assert (gmin[:-1] >= gmax[1:]).all()                  # ... by decreasing valence from group to group
        s = slots[ts:ts + n].ravel().astype(int); l = conn[ts:ts + n].ravel().astype(int)
        assert len(set(s.tolist())) == 4 * n                 # every corner owns exactly one slot
        rank = np.empty(nv, int); rank[perm] = np.arange(nv)  # local id -> reduce-order position
        t = rank[l]
        assert ((s >= start[t]) & (s < start[t] + cnt[t])).all()   # ... inside its vertex's range
        assert np.array_equal(np.bincount(l, minlength=nv)[perm], cnt)


def test_tables_keep_shared_memory_accesses_nearly_conflict_free(native_lib):
    """tools/smem_model.py replays the kernels' warp-wide shared-memory accesses on the packed tables
    (16-byte accesses per quarter warp, 8-byte per half warp).  The local ids, the reduce order / pads and
    the slot positions are chosen by tiling.cpp to avoid bank conflicts: guard the achieved level."""
    import sys

    sys.path.insert(0, str(ROOT / "tools"))
    import smem_model

    tiles, conn, slots, tv, voff, vperm = smem_model.host_tables(20)
    per, ideal = smem_model.model(tiles, conn, slots, voff, vperm, quarter=True, max_tiles=40)
    assert per["gather"] <= 1.2 * ideal["gather"]            # ascending local ids: 1.8x
    got, asc = smem_model.global_lines(tiles, tv, 40)
    assert got <= 1.001 * asc                                # every warp still gathers an ascending window of 32
    assert per["red_ld"] <= 1.10 * ideal["red_ld"]           # odd-length padding only: 1.7x
    assert per["slot_st"] <= 1.45 * ideal["slot_st"]         # tet-order greedy: 1.7x


def test_tiling_splits_on_vertex_budget(native_lib):
    """Disconnected tets: 4 new vertices each, so a tile closes at 64 tets (256 vertices)."""
    from apple_b200.mesh import TetMesh

    n = 300
    pts = np.random.default_rng(0).random((4 * n, 3))
    mesh = TetMesh(pts, np.arange(4 * n).reshape(n, 4))
    X = mesh.points[mesh.cells]
    vol = np.einsum("ci,ci->c", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 3] - X[:, 0])
    mesh.cells[vol < 0] = mesh.cells[vol < 0][:, [0, 2, 1, 3]]
    tiles, *_ = _host_tables(native_lib, mesh, with_points=False)
    limit = tiles[:, 3].max()
    assert limit <= 256 and (tiles[:-1, 3] == limit).all() and (tiles[:-1, 1] == limit // 4).all()


def test_setup_rejects_bad_meshes(native_lib):
    from apple_b200 import _lib
    from oracle import region

    mesh, _, _ = make_case(n=2, seed=0)
    dhdX, dV = region.compute_grad(mesh.points, mesh.cells)
    T, V = mesh.n_cells, mesh.n_points
    ones = np.ones(T)
    P = _lib.host_ptr
    h = ctypes.c_void_p()
    bad = mesh.cells.copy(); bad[3, 2] = V
    assert native_lib.apl_fem_create(0, 1, T, V, P(bad), P(dhdX), P(dV), P(ones), P(ones), None, None, -1,
                                     ctypes.byref(h)) == -3
    assert b"outside" in native_lib.apl_last_error()
    d2 = dhdX.copy(); d2[5, 0, 1] += 1.0
    assert native_lib.apl_fem_create(0, 1, T, V, P(mesh.cells), P(d2), P(dV), P(ones), P(ones), None, None, -1,
                                     ctypes.byref(h)) == -3
    assert b"sum to zero" in native_lib.apl_last_error()
    assert native_lib.apl_fem_create(2, 1, T, V, P(mesh.cells), P(dhdX), P(dV), P(ones), P(ones), None, None, -1,
                                     ctypes.byref(h)) == -1  # muscle without activation


@pytest.fixture(scope="module")
def host_math(tmp_path_factory):
    out = tmp_path_factory.mktemp("native") / "elem_host.so"
    src = ROOT / "tests" / "native" / "elem_host.cpp"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(out), str(src)], check=True)
    return ctypes.CDLL(str(out))


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-11), (np.float32, 2e-5)], ids=["f64", "f32"])
@pytest.mark.parametrize("kind", KINDS)
def test_per_tet_math_header_matches_oracle(host_math, kind, dtype, tol):
    """apple_b200/csrc/elem_math.cuh (the code the kernels inline) == oracle, element by element."""
    mesh, u, p = make_case(n=5, seed=2, amp=0.15)
    ora = oracle_potential(kind, mesh)
    T = mesh.n_cells
    nrec = 18 if kind == "muscle" else 12
    rec = np.zeros((T, nrec), dtype)
    rec[:, :9] = ora.dhdX[:, 1:4].reshape(T, 9)
    rec[:, 9] = ora.dV
    rec[:, 10] = ora.materials["mu"]
    if kind != "arap":
        rec[:, 11] = ora.materials["lambda_"]
    if kind == "muscle":
        rec[:, 12:] = ora.materials["activation"]
    uc = np.ascontiguousarray(u[mesh.cells], dtype); pc = np.ascontiguousarray(p[mesh.cells], dtype)
    psi = np.zeros(T, dtype); quad = np.zeros(T, dtype)
    g, dg, hp = (np.zeros((T, 4, 3), dtype) for _ in range(3))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    host_math.elem_eval_host(KINDS.index(kind), int(dtype == np.float64), T, P(rec), P(uc), P(pc), P(psi), P(quad),
                             P(g), P(dg), P(hp))

    def rel(a, b):
        a = a.reshape(T, -1).astype(np.float64); b = b.reshape(T, -1)
        return (np.abs(a - b).max(1) / np.abs(b).max(1)).max()

    assert rel(g, ora.elem_grad(u)) < tol
    assert rel(dg, ora.elem_hess_diag(u)) < tol
    assert rel(hp, ora.elem_hess_prod(u, p)) < tol
    e, q = ora.elem_fun(u), ora.elem_hess_quad(u, p)
    assert np.abs(psi - e).max() < tol * np.abs(e).max() * 10
    assert np.abs(quad - q).max() < tol * np.abs(q).max() * 10


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 5e-6)], ids=["f64", "f32"])
def test_svd3_rotation_variant_convention(host_math, dtype, tol):
    """U, V proper rotations, s0 >= s1 >= |s2|, sign(s2) = sign(det F), F == U diag(s) V^T."""
    rng = np.random.default_rng(3)
    n = 4000
    F = np.eye(3)[None] + 0.5 * rng.standard_normal((n, 3, 3))
    F[:50] = np.eye(3)                                     # repeated singular values
    F[50:100, :, 2] *= 1e-3                                # nearly flat
    F = np.ascontiguousarray(F, dtype)
    U = np.zeros_like(F); V = np.zeros_like(F); s = np.zeros((n, 3), dtype)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    host_math.svd3_host(int(dtype == np.float64), n, P(F), P(U), P(s), P(V))
    U64, V64, s64 = U.astype(float), V.astype(float), s.astype(float)
    np.testing.assert_allclose(np.linalg.det(U64), 1.0, atol=10 * tol)
    np.testing.assert_allclose(np.linalg.det(V64), 1.0, atol=10 * tol)
    rec = np.einsum("nij,nj,nkj->nik", U64, s64, V64)
    assert np.abs(rec - F).max() < 20 * tol
    assert (s64[:, 0] >= s64[:, 1] - tol).all() and (s64[:, 1] >= np.abs(s64[:, 2]) - tol).all()
    assert (np.sign(s64[:, 2]) == np.sign(np.linalg.det(F.astype(float))))[100:].all()
    ref = np.linalg.svd(F.astype(float), compute_uv=False)
    assert np.abs(np.abs(s64) - ref).max() < 20 * tol


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-6)], ids=["f64", "f32"])
def test_arap_closed_form_polar_and_twist_operator(host_math, dtype, tol):
    """polar_twist (no iterative SVD: invariants of S from the trigonometric eigenvalues of F^T F,
    R = (I1 F - F S + cof F) / I2, Lam = f(tr(S) - S) by Newton interpolation) against numpy's SVD in the
    rotation-variant convention of warp/math/_rotation.py:9-22 and the clamped twist rates of
    warp/fem/func/_misc.py:31-43 -- rest state, isotropic scaling, small / large strain, flat and inverted
    elements (which take the Jacobi fallback), and agreement of the two paths."""
    rng = np.random.default_rng(11)
    n = 3000
    Q, _ = np.linalg.qr(rng.standard_normal((n, 3, 3)))
    Q[np.linalg.det(Q) < 0, :, 0] *= -1
    cases = {
        "identity": np.repeat(np.eye(3)[None], 64, 0),
        "isotropic": np.eye(3)[None] * rng.uniform(0.5, 2.0, (256, 1, 1)),
        "strain 1e-4": np.eye(3) + 1e-4 * rng.standard_normal((n, 3, 3)),
        "strain 3e-2": Q @ (np.eye(3) + 3e-2 * rng.standard_normal((n, 3, 3))),
        "strain 0.2": Q @ (np.eye(3) + 0.2 * rng.standard_normal((n, 3, 3))),
        "flat": (np.eye(3) + 0.1 * rng.standard_normal((n, 3, 3))) @ np.diag([1.0, 1.0, 1e-3]),
        "inverted": (np.eye(3) + 0.2 * rng.standard_normal((n, 3, 3))) @ np.diag([1.0, 1.0, -0.3]),
    }
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    for name, F64 in cases.items():
        F = np.ascontiguousarray(F64, dtype)
        Ft = F.astype(np.float64)
        U, s, Vt = np.linalg.svd(Ft)
        flip = np.linalg.det(U) < 0; U[flip, :, 2] *= -1; s[flip, 2] *= -1
        flip = np.linalg.det(Vt) < 0; Vt[flip, 2, :] *= -1; s[flip, 2] *= -1
        R_ref = U @ Vt
        lam = 2.0 / np.maximum(np.stack([s[:, 1] + s[:, 2], s[:, 0] + s[:, 2], s[:, 0] + s[:, 1]], 1), 2.0)
        L_ref = np.einsum("nki,nk,nkj->nij", Vt, lam, Vt)          # sum_k lam_k v_k v_k^T, axis k <-> s_i + s_j, i,j != k
        out = {}
        for path in (0, 1):
            R = np.zeros_like(F); L = np.zeros((len(F), 6), dtype); sg = np.zeros((len(F), 3), dtype)
            host_math.polar_twist_host(int(dtype == np.float64), path, len(F), P(F), P(R), P(L), P(sg))
            out[path] = (R.astype(float), L.astype(float), sg.astype(float))
        # well-posed elements: the two smallest singular values do not cancel (R is unique)
        ok = (s[:, 1] + s[:, 2]) > 0.2 * s[:, 0]
        assert ok.mean() > 0.7, name
        for path, (R, L, sg) in out.items():
            Lm = np.stack([L[:, [0, 3, 4]], L[:, [3, 1, 5]], L[:, [4, 5, 2]]], 1)
            scale = s[:, 0][ok]
            assert (np.abs(R - R_ref).max((1, 2))[ok] < 40 * tol).all(), (name, path)
            assert (np.abs(Lm - L_ref).max((1, 2))[ok] < 40 * tol).all(), (name, path)
            assert (np.abs(np.sort(sg, 1) - np.sort(s, 1)).max(1)[ok] < 400 * tol * scale).all(), (name, path)
            assert np.isfinite(R).all() and np.isfinite(L).all()
        if name in ("strain 3e-2", "strain 1e-4", "identity", "isotropic"):
            # the regime of elasticity solves: a few ulps
            assert np.abs(out[0][0] - R_ref).max() < 4 * tol and np.abs(out[0][1] - out[1][1]).max() < 8 * tol, name


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("kind", ["snh", "arap", "muscle", "snh+arap"])
def test_static_planes_hold_the_reference_arrays(native_lib, kind, dtype):
    """The packed planes the kernels stream are exactly the reference's region / material arrays:
    rows 1..3 of dhdX, Fraction * dV, mu, lambda, activation -- in packed (tile) order."""
    from apple_b200 import _lib
    from oracle import region

    mesh, _, _ = make_case(n=5, seed=3)
    T, V = mesh.n_cells, mesh.n_points
    dhdX, dV = region.compute_grad(mesh.points, mesh.cells, mesh.cell_data["Fraction"], dtype=dtype)
    mu = mesh.cell_data["mu"].astype(dtype); la = mesh.cell_data["lambda"].astype(dtype)
    act = mesh.cell_data["activation"].astype(dtype)
    dV2 = (0.5 * dV).astype(dtype); mu2 = (2.0 * mu).astype(dtype)
    P = _lib.host_ptr
    h = ctypes.c_void_p()
    code = _lib.F32 if dtype == np.float32 else _lib.F64
    cells = np.ascontiguousarray(mesh.cells, dtype=np.int32); pts = np.ascontiguousarray(mesh.points)
    if kind == "snh+arap":
        rc = native_lib.apl_fem_create_snh_arap(code, T, V, P(cells), P(dhdX), P(dV), P(mu), P(la), P(dV2), P(mu2), P(pts),
                                                -1, ctypes.byref(h))
    else:
        k = {"snh": 0, "arap": 1, "muscle": 2}[kind]
        rc = native_lib.apl_fem_create(k, code, T, V, P(cells), P(dhdX), P(dV), P(mu), P(la), P(act), P(pts), -1,
                                       ctypes.byref(h))
    assert rc == 0, native_lib.apl_last_error()
    npl, stride = ctypes.c_int64(), ctypes.c_int64()
    native_lib.apl_fem_host_planes(h, None, ctypes.byref(npl), ctypes.byref(stride))
    vec = 16 // np.dtype(dtype).itemsize
    planes = np.zeros((npl.value, stride.value, vec), dtype)
    native_lib.apl_fem_host_planes(h, P(planes), None, None)
    order = np.zeros(T, np.int64)
    native_lib.apl_fem_host_tables(h, None, P(order), None, None, None, None, None)
    native_lib.apl_fem_destroy(h)
    rec = planes.transpose(1, 0, 2).reshape(stride.value, npl.value * vec)[:T]      # (T, padded record)
    np.testing.assert_array_equal(rec[:, :9], dhdX[order][:, 1:4].reshape(T, 9))
    np.testing.assert_array_equal(rec[:, 9], dV[order])
    np.testing.assert_array_equal(rec[:, 10], mu[order])
    if kind != "arap":
        np.testing.assert_array_equal(rec[:, 11], la[order])
    if kind == "muscle":
        np.testing.assert_array_equal(rec[:, 12:18], act[order])
    if kind == "snh+arap":
        np.testing.assert_array_equal(rec[:, 12], dV2[order])
        np.testing.assert_array_equal(rec[:, 13], mu2[order])
    assert npl.value == -(-{"snh": 12, "arap": 12, "muscle": 18, "snh+arap": 14}[kind] // vec)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-11), (np.float32, 3e-5)], ids=["f64", "f32"])
def test_fused_record_equals_sum_of_two_potentials(host_math, dtype, tol):
    """KIND SNH_ARAP (two potentials of one cell in one evaluation) == SNH + ARAP of the oracle, with
    per-potential volumes/materials and per-potential clamps."""
    from oracle import fem as ofem

    mesh, u, p = make_case(n=4, seed=8, amp=0.6)   # large deformation: some clamps are active
    m2 = mesh.copy()
    m2.cell_data["Fraction"] = 1.0 - 0.5 * mesh.cell_data["Fraction"]
    m2.cell_data["mu"] = mesh.cell_data["mu"][::-1].copy()
    a, b = oracle_potential("snh", mesh), oracle_potential("arap", m2)
    T = mesh.n_cells
    rec = np.zeros((T, 14), dtype)
    rec[:, :9] = a.dhdX[:, 1:4].reshape(T, 9)
    rec[:, 9], rec[:, 10], rec[:, 11] = a.dV, a.materials["mu"], a.materials["lambda_"]
    rec[:, 12], rec[:, 13] = b.dV, b.materials["mu"]
    uc = np.ascontiguousarray(u[mesh.cells], dtype); pc = np.ascontiguousarray(p[mesh.cells], dtype)
    psi = np.zeros(T, dtype); quad = np.zeros(T, dtype)
    g, dg, hp = (np.zeros((T, 4, 3), dtype) for _ in range(3))
    P = lambda x: x.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    host_math.elem_eval_host(3, int(dtype == np.float64), T, P(rec), P(uc), P(pc), P(psi), P(quad), P(g), P(dg), P(hp))

    def rel(x, y):
        x = x.reshape(T, -1).astype(np.float64); y = y.reshape(T, -1)
        return (np.abs(x - y).max(1) / np.abs(y).max(1)).max()

    assert rel(g, a.elem_grad(u) + b.elem_grad(u)) < tol
    assert rel(dg, a.elem_hess_diag(u) + b.elem_hess_diag(u)) < tol
    assert rel(hp, a.elem_hess_prod(u, p) + b.elem_hess_prod(u, p)) < tol
    q = a.elem_hess_quad(u, p) + b.elem_hess_quad(u, p)
    e = a.elem_fun(u) + b.elem_fun(u)
    assert np.abs(quad - q).max() < 10 * tol * np.abs(q).max()
    assert np.abs(psi - e).max() < 10 * tol * np.abs(e).max()
    qa = a.elem_hess_quad(u, p); a.clamp_hess_quad = False
    assert (a.elem_hess_quad(u, p) < 0).any() and (qa >= 0).all()    # the per-potential clamp is exercised


def test_tile_tables_assemble_the_oracle_gradient(native_lib):
    """Emulates, in numpy, exactly how the element kernels interpret the tables (gather through conn /
    tile_verts, one slot per corner, per-vertex slot ranges from tile_voff in tile_vperm order with the
    pad count in bits 12..15, flush to tile_verts) and checks that the assembled field is the oracle's."""
    mesh, u, _ = make_case(n=6, seed=5, morton=False)
    ora = oracle_potential("snh", mesh)
    tiles, order, conn, slots, tv, voff, vperm = _host_tables(native_lib, mesh)
    elem = ora.elem_grad(u)[order]                      # per-corner contributions in packed order
    uq = u[mesh.cells[order]]                           # what a gather through the global ids must return
    out = np.zeros_like(u)
    for (ts, n, vs, nv, vo, nslots) in tiles:
        verts = tv[vs:vs + nv]
        c = conn[ts:ts + n].astype(int)
        assert np.array_equal(u[verts][c], uq[ts:ts + n])          # phase 2: shared-memory gather
        buf = np.full((nslots, 3), np.nan)
        buf[slots[ts:ts + n].astype(int).ravel()] = elem[ts:ts + n].reshape(-1, 3)   # phase 3: slot stores
        raw = voff[vo:vo + nv + 1].astype(int)
        start, padded = raw & 0x0fff, raw[:-1] >> 12
        assert start[-1] == nslots
        acc = np.zeros((nv, 3))
        for t in range(nv):                                        # phase 4: reduce in tile_vperm order
            cnt = start[t + 1] - start[t] - padded[t]
            rows = buf[start[t]:start[t] + cnt]
            assert not np.isnan(rows).any()                        # only real slots are read
            acc[vperm[vs + t]] = rows.sum(axis=0)
        np.add.at(out, verts, acc)                                 # phase 5: one RED per tile vertex
    ref = np.zeros_like(u); ora.grad(u, ref)
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())


def test_set_materials_patches_only_the_requested_columns(native_lib):
    """apl_fem_set_materials (replaces re-creating the Materials struct, warp/fem/utils/_material.py:15-31):
    on a host-only handle the patched planes equal those of a handle created with the new arrays."""
    from apple_b200 import _lib
    from oracle import region

    mesh, _, _ = make_case(n=4, seed=12)
    T, V = mesh.n_cells, mesh.n_points
    dhdX, dV = region.compute_grad(mesh.points, mesh.cells, mesh.cell_data["Fraction"])
    mu, la, act = mesh.cell_data["mu"], mesh.cell_data["lambda"], mesh.cell_data["activation"]
    rng = np.random.default_rng(0)
    act2, mu2 = 0.2 * rng.standard_normal((T, 6)), mu * 3.0
    P = _lib.host_ptr
    cells = np.ascontiguousarray(mesh.cells, dtype=np.int32); pts = np.ascontiguousarray(mesh.points)

    def create(mu_, act_):
        h = ctypes.c_void_p()
        assert native_lib.apl_fem_create(2, _lib.F64, T, V, P(cells), P(dhdX), P(dV), P(np.ascontiguousarray(mu_)), P(la),
                                         P(np.ascontiguousarray(act_)), P(pts), -1, ctypes.byref(h)) == 0
        return h

    def planes(h):
        npl, stride = ctypes.c_int64(), ctypes.c_int64()
        native_lib.apl_fem_host_planes(h, None, ctypes.byref(npl), ctypes.byref(stride))
        out = np.zeros((npl.value, stride.value, 2))
        native_lib.apl_fem_host_planes(h, P(out), None, None)
        return out

    a, b = create(mu, act), create(mu2, act2)
    before = planes(a)
    assert native_lib.apl_fem_set_materials(a, None, P(np.ascontiguousarray(mu2)), None, P(np.ascontiguousarray(act2))) == 0
    after = planes(a)
    np.testing.assert_array_equal(after, planes(b))
    assert not np.array_equal(before, after)
    native_lib.apl_fem_destroy(a); native_lib.apl_fem_destroy(b)


def test_mark_boundary_partitions_the_tile_headers(native_lib):
    """apl_fem_mark_boundary (multi-GPU overlap): tiles touching a flagged vertex move to the front of the
    header list, nothing else changes -- the interior part touches no flagged vertex, the headers are a
    permutation of the original ones, NULL flags restore a single part."""
    from apple_b200 import _lib
    from oracle import region

    mesh, _, _ = make_case(n=9, seed=0, morton=True)
    dhdX, dV = region.compute_grad(mesh.points, mesh.cells)
    T, V = mesh.n_cells, mesh.n_points
    one = np.ones(T)
    P = _lib.host_ptr
    h = ctypes.c_void_p()
    cells = np.ascontiguousarray(mesh.cells, dtype=np.int32); pts = np.ascontiguousarray(mesh.points)
    assert native_lib.apl_fem_create(0, _lib.F64, T, V, P(cells), P(dhdX), P(dV), P(one), P(one), None, P(pts), -1,
                                     ctypes.byref(h)) == 0
    info = (ctypes.c_int64 * 10)(); native_lib.apl_fem_info(h, info)
    nt, nv = info[2], info[3]

    def tables():
        tiles = np.zeros((nt, 6), np.int32); tv = np.zeros(nv, np.int32)
        native_lib.apl_fem_host_tables(h, P(tiles), None, None, None, P(tv), None, None)
        return tiles, tv

    before, tv = tables()
    flags = np.zeros(V, np.uint8); flags[mesh.points[:, 0] == 0.0] = 1      # one face of the cube is "shared"
    nb = ctypes.c_int64()
    assert native_lib.apl_fem_mark_boundary(h, P(flags), ctypes.byref(nb)) == 0
    after, tv2 = tables()
    assert np.array_equal(tv, tv2) and 0 < nb.value < nt
    touches = np.array([flags[tv[vs:vs + n]].any() for (_, _, vs, n, _, _) in after])
    assert touches[:nb.value].all() and not touches[nb.value:].any()
    assert sorted(map(tuple, before.tolist())) == sorted(map(tuple, after.tolist()))
    # stable: each part keeps the packed (Morton) order
    assert (np.diff(after[:nb.value, 0]) > 0).all() and (np.diff(after[nb.value:, 0]) > 0).all()
    assert native_lib.apl_fem_mark_boundary(h, None, ctypes.byref(nb)) == 0 and nb.value == 0
    assert np.array_equal(tables()[0], before)
    assert native_lib.apl_fem_eval_part(h, 3, 1, None, None, 3, None, None, None, None, None, 3, 0, None) != 0  # bad part
    native_lib.apl_fem_destroy(h)


@pytest.fixture(scope="module")
def tile_host(tmp_path_factory):
    out = tmp_path_factory.mktemp("native_tile") / "tile_host.so"
    src = ROOT / "tests" / "native" / "tile_host.cpp"
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(out), str(src)], check=True)
    return ctypes.CDLL(str(out))


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-11), (np.float32, 1e-5)], ids=["f64", "f32"])
@pytest.mark.parametrize("kind", ["snh", "arap", "muscle", "snh+arap"])
def test_consumer_logic_replayed_on_the_host_matches_oracle(native_lib, tile_host, kind, dtype, tol):
    """The consumer-side DEVICE code of the element kernels (csrc/tile_logic.cuh: record / connectivity decoding,
    corner gather, slot stores, per-lane slot reduction with the tile_voff decoding, direct flush)
    compiled for the host and replayed thread by thread on the packed tables and planes of a host-only handle ==
    the oracle's assembled energy / gradient / diagonal / HVP / quadratic form."""
    from apple_b200 import _lib
    from oracle import fem as ofem
    from oracle import region

    mesh, u, p = make_case(n=6, seed=9, amp=0.12)
    T, V = mesh.n_cells, mesh.n_points
    dhdX, dV = region.compute_grad(mesh.points, mesh.cells, mesh.cell_data["Fraction"], dtype=dtype)
    mu, la = mesh.cell_data["mu"].astype(dtype), mesh.cell_data["lambda"].astype(dtype)
    act = mesh.cell_data["activation"].astype(dtype)
    m2 = mesh.copy()
    m2.cell_data["Fraction"] = 1.0 - 0.5 * mesh.cell_data["Fraction"]
    m2.cell_data["mu"] = mesh.cell_data["mu"][::-1].copy()
    if kind == "snh+arap":
        ora = ofem.Model([oracle_potential("snh", mesh), oracle_potential("arap", m2)], V)
        dV2 = region.compute_grad(m2.points, m2.cells, m2.cell_data["Fraction"], dtype=dtype)[1]
        mu2 = m2.cell_data["mu"].astype(dtype)
    else:
        ora = ofem.Model([oracle_potential(kind, mesh)], V)
    P = _lib.host_ptr
    cells = np.ascontiguousarray(mesh.cells, dtype=np.int32); pts = np.ascontiguousarray(mesh.points)
    code = _lib.F32 if dtype == np.float32 else _lib.F64
    h = ctypes.c_void_p()
    if kind == "snh+arap":
        rc = native_lib.apl_fem_create_snh_arap(code, T, V, P(cells), P(dhdX), P(dV), P(mu), P(la), P(dV2), P(mu2),
                                                P(pts), -1, ctypes.byref(h))
    else:
        k = {"snh": 0, "arap": 1, "muscle": 2}[kind]
        rc = native_lib.apl_fem_create(k, code, T, V, P(cells), P(dhdX), P(dV), P(mu), P(la), P(act), P(pts), -1,
                                       ctypes.byref(h))
    assert rc == 0, native_lib.apl_last_error()
    info = (ctypes.c_int64 * 10)(); native_lib.apl_fem_info(h, info)
    nt, nv, nvo, npk = info[2], info[3], info[8], info[9]
    rows, width = T, 4
    tiles = np.zeros((nt, 6), np.int32); conn = np.zeros((rows, width), np.uint8); slots = np.zeros((rows, width), np.uint16)
    tv = np.zeros(nv, np.int32); voff = np.zeros(nvo, np.uint16); vperm = np.zeros(nv, np.uint8)
    native_lib.apl_fem_host_tables(h, P(tiles), None, P(conn), P(slots), P(tv), P(voff), P(vperm))
    npl, stride = ctypes.c_int64(), ctypes.c_int64()
    native_lib.apl_fem_host_planes(h, None, ctypes.byref(npl), ctypes.byref(stride))
    planes = np.zeros((npl.value, stride.value, 16 // np.dtype(dtype).itemsize), dtype)
    native_lib.apl_fem_host_planes(h, P(planes), None, None)
    native_lib.apl_fem_destroy(h)

    ud, pd = np.ascontiguousarray(u, dtype), np.ascontiguousarray(p, dtype)
    ref = {"fun": ora.fun(u), "quad": ora.hess_quad(u, p), "grad": ora.grad(u), "diag": ora.hess_diag(u),
           "prod": ora.hess_prod(u, p)}
    kcode = {"snh": 0, "arap": 1, "muscle": 2, "snh+arap": 3}[kind]
    tile_host.tile_emulate.argtypes = [ctypes.c_int] * 3 + [ctypes.c_int64] + [ctypes.c_void_p] * 7 + [ctypes.c_int64] + \
        [ctypes.c_void_p] * 2 + [ctypes.c_double] + [ctypes.c_void_p] * 5
    for ops in (11, 7, 16, 15):
        grad, diag, prod = (np.zeros((V, 3), dtype) for _ in range(3))
        fun, quad = np.zeros(1), np.zeros(1)
        rc = tile_host.tile_emulate(kcode, int(dtype == np.float64), ops, nt, P(tiles), P(conn), P(slots), P(tv),
                                    P(voff), P(vperm), P(planes), stride.value, P(ud), P(pd), 0.0, P(grad), P(diag),
                                    P(prod), P(fun), P(quad))
        assert rc == 0
        got = {"fun": fun[0], "quad": quad[0], "grad": grad, "diag": diag, "prod": prod}
        for bit, name in ((1, "fun"), (16, "quad"), (2, "grad"), (4, "diag"), (8, "prod")):
            if ops & bit:
                a, b = np.asarray(got[name], np.float64), np.asarray(ref[name])
                assert np.isfinite(a).all(), (name, ops)                      # a NaN = a slot read before it was written
                assert np.abs(a - b).max() <= tol * np.abs(b).max(), (kind, ops, name)
    # PNCG's trial pass: energy / gradient / diagonal at u + alpha p, the trial point formed inside the gather
    alpha = 0.01
    grad, diag, prod = (np.zeros((V, 3), dtype) for _ in range(3))
    fun, quad = np.zeros(1), np.zeros(1)
    assert tile_host.tile_emulate(kcode, int(dtype == np.float64), 7, nt, P(tiles), P(conn), P(slots), P(tv), P(voff),
                                  P(vperm), P(planes), stride.value, P(ud), P(pd), alpha, P(grad), P(diag), P(prod), P(fun),
                                  P(quad)) == 0
    ut = u + alpha * p
    for a, b in ((fun[0], ora.fun(ut)), (grad, ora.grad(ut)), (diag, ora.hess_diag(ut))):
        assert np.abs(np.asarray(a, np.float64) - b).max() <= 3 * tol * np.abs(b).max(), (kind, "axpy")


def test_tiling_is_deterministic_across_host_thread_counts(native_lib, monkeypatch):
    """Tiles are packed in parallel (one scratch per host thread, per-tile random seeds): the tables must not depend
    on the number of threads -- every rank of a sharded run and every rerun sees the same layout."""
    mesh, _, _ = make_case(n=24, seed=0)          # 270 tiles: above the threshold of the parallel path
    out = []
    for threads in ("1", "5"):
        monkeypatch.setenv("APL_TILING_THREADS", threads)
        out.append(_host_tables(native_lib, mesh))
    for a, b in zip(*out):
        assert np.array_equal(a, b)


def test_unstructured_delaunay_mesh_through_the_host_replay(native_lib, tile_host):
    """An unstructured mesh (Delaunay tetrahedralisation of random points, cells in arbitrary order, uneven
    valences): tiles close on the vertex budget as well as on the tet count -- the replayed consumer logic must still assemble the oracle's energy / gradient / HVP."""
    scipy_spatial = pytest.importorskip("scipy.spatial")
    from apple_b200 import _lib
    from apple_b200.mesh import TetMesh
    from oracle import fem as ofem
    from oracle import region

    rng = np.random.default_rng(1)
    pts = rng.random((1500, 3))
    cells = scipy_spatial.Delaunay(pts).simplices.astype(np.int32)
    X = pts[cells]
    vol = np.einsum("ci,ci->c", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 3] - X[:, 0])
    cells[vol < 0] = cells[vol < 0][:, [0, 2, 1, 3]]
    cells = np.ascontiguousarray(cells[np.abs(vol) > 1e-6])          # drop slivers (their dhdX is huge)
    mesh = TetMesh(pts, cells)
    T, V = mesh.n_cells, mesh.n_points
    dhdX, dV = region.compute_grad(mesh.points, mesh.cells)
    mu, la = rng.uniform(1, 3, T), rng.uniform(1, 9, T)
    u, p = 0.003 * rng.standard_normal((V, 3)), rng.standard_normal((V, 3))
    ora = ofem.Model([ofem.StableNeoHookean(mesh.cells, dhdX, dV, mu=mu, lambda_=la)], V)
    P = _lib.host_ptr
    h = ctypes.c_void_p()
    rc = native_lib.apl_fem_create(0, _lib.F64, T, V, P(cells), P(dhdX), P(dV), P(mu), P(la), None,
                                   P(np.ascontiguousarray(pts)), -1, ctypes.byref(h))
    assert rc == 0, native_lib.apl_last_error()
    info = (ctypes.c_int64 * 10)(); native_lib.apl_fem_info(h, info)
    nt, nv, nvo, npk = info[2], info[3], info[8], info[9]
    rows, width = T, 4
    tiles = np.zeros((nt, 6), np.int32); conn = np.zeros((rows, width), np.uint8); slots = np.zeros((rows, width), np.uint16)
    tv = np.zeros(nv, np.int32); voff = np.zeros(nvo, np.uint16); vperm = np.zeros(nv, np.uint8)
    native_lib.apl_fem_host_tables(h, P(tiles), None, P(conn), P(slots), P(tv), P(voff), P(vperm))
    npl, stride = ctypes.c_int64(), ctypes.c_int64()
    native_lib.apl_fem_host_planes(h, None, ctypes.byref(npl), ctypes.byref(stride))
    planes = np.zeros((npl.value, stride.value, 2)); native_lib.apl_fem_host_planes(h, P(planes), None, None)
    native_lib.apl_fem_destroy(h)
    assert (tiles[:, 3] <= 192).all() and (tiles[:, 1] <= 256).all() and tiles[:, 1].sum() == npk
    grad, diag, prod = (np.zeros((V, 3)) for _ in range(3))
    fun, quad = np.zeros(1), np.zeros(1)
    tile_host.tile_emulate.argtypes = [ctypes.c_int] * 3 + [ctypes.c_int64] + [ctypes.c_void_p] * 7 + [ctypes.c_int64] + \
        [ctypes.c_void_p] * 2 + [ctypes.c_double] + [ctypes.c_void_p] * 5
    assert tile_host.tile_emulate(0, 1, 11, nt, P(tiles), P(conn), P(slots), P(tv), P(voff), P(vperm), P(planes),
                                  stride.value, P(u), P(p), 0.0, P(grad), P(diag), P(prod), P(fun), P(quad)) == 0
    for a, b in ((fun[0], ora.fun(u)), (grad, ora.grad(u)), (prod, ora.hess_prod(u, p))):
        assert np.abs(np.asarray(a) - b).max() <= 1e-11 * np.abs(b).max()


@pytest.mark.parametrize("kind", KINDS)
def test_mixed_derivative_closed_forms_match_finite_differences_of_the_oracle(host_math, kind):
    """elem_mixed (csrc/elem_math.cuh): d/dq [grad E . p] per cell for q in (mu, lambda, activation) -- the
    reference's historical mixed_derivative_prod -- against central differences of the oracle's per-cell gradient
    with respect to its material arrays."""
    import copy

    mesh, u, p = make_case(n=4, seed=2, amp=0.15)
    T = mesh.n_cells
    ora = oracle_potential(kind, mesh)
    nrec = 18 if kind == "muscle" else 12
    rec = np.zeros((T, nrec))
    rec[:, :9] = ora.dhdX[:, 1:4].reshape(T, 9); rec[:, 9] = ora.dV; rec[:, 10] = ora.materials["mu"]
    if kind != "arap":
        rec[:, 11] = ora.materials["lambda_"]
    if kind == "muscle":
        rec[:, 12:] = ora.materials["activation"]
    uc = np.ascontiguousarray(u[mesh.cells]); pc = np.ascontiguousarray(p[mesh.cells])
    dm, dl, da = np.zeros(T), np.zeros(T), np.zeros((T, 6))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    host_math.elem_mixed_host(KINDS.index(kind), 1, T, P(rec), P(uc), P(pc), P(dm), P(dl), P(da))

    def phi(pot):                                   # per-cell  grad E . p
        return np.einsum("cai,cai->c", pot.elem_grad(u), p[mesh.cells])

    def fd(name, comp=None, eps=1e-6):
        lo, hi = copy.deepcopy(ora), copy.deepcopy(ora)
        for pot, sign in ((hi, 1.0), (lo, -1.0)):
            m = dict(ora.materials)
            x = np.array(ora.materials[name], dtype=float, copy=True)
            if comp is None:
                x += sign * eps
            else:
                x[:, comp] += sign * eps
            m[name] = x
            pot.materials = m
        return (phi(hi) - phi(lo)) / (2 * eps)

    rel = lambda got, ref: np.abs(got - ref).max() / np.abs(ref).max()  # noqa: E731
    assert rel(dm, fd("mu")) < 1e-6
    if kind != "arap":
        assert rel(dl, fd("lambda_")) < 1e-6
    if kind == "muscle":
        for k in range(6):
            assert rel(da[:, k], fd("activation", k)) < 1e-6
