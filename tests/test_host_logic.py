"""CPU: host-side logic that needs no GPU -- mesh generators, Morton reordering, attribute names,
DOF maps (torch CPU tensors), termination classification."""

import numpy as np
import torch

from apple_b200.common import ACTIVATION, FIXED_MASK, GLOBAL_POINT_ID, LAMBDA, MU, lame_converter
from apple_b200.forward.dof_map import DofMap, DofMapBuilder
from apple_b200.mesh import TetMesh, cube_tet_mesh, embedded_tetra_mesh, lumped_vertex_volume, morton_reorder
from apple_b200.optim import Result
from apple_b200.optim._pncg import ConvergenceCriteria, _classify


def _volumes(mesh):
    X = mesh.points[mesh.cells]
    return np.einsum("ci,ci->c", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 3] - X[:, 0]) / 6.0


def test_cube_mesh_is_conforming_and_positive():
    mesh = cube_tet_mesh(6, grading=1.05)
    assert mesh.n_cells == 5 * 6 ** 3 and mesh.n_points == 7 ** 3
    vol = _volumes(mesh)
    assert (vol > 0).all() and abs(vol.sum() - 1.0) < 1e-12
    faces = np.sort(mesh.cells[:, [[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]]].reshape(-1, 3), axis=1)
    _, counts = np.unique(faces, axis=0, return_counts=True)
    assert set(counts.tolist()) <= {1, 2}                       # conforming: a face has 1 or 2 tets
    assert (counts == 1).sum() == 6 * 2 * 6 * 6                 # boundary triangles
    assert abs(lumped_vertex_volume(mesh).sum() - 1.0) < 1e-12  # benches/test_aggregation.py:82,129


def test_config_sizes_of_the_baseline():
    assert cube_tet_mesh(16, morton=False).n_cells == 20_480     # C1
    assert 5 * 58 ** 3 == 975_560 and 59 ** 3 == 205_379         # C2 (not built here: size only)


def test_morton_reorder_is_a_relabelling():
    mesh = cube_tet_mesh(4, morton=False)
    mesh.cell_data["mu"] = np.arange(mesh.n_cells, dtype=float)
    mesh.point_data["tag"] = np.arange(mesh.n_points)
    re = morton_reorder(mesh)
    assert np.allclose(np.sort(_volumes(re)), np.sort(_volumes(mesh)))
    # cell data follows its cell: compare centroids keyed by the tag
    c0 = {int(k): v for k, v in zip(mesh.cell_data["mu"], mesh.points[mesh.cells].mean(1).round(9).tolist())}
    for k, c in zip(re.cell_data["mu"], re.points[re.cells].mean(1).round(9).tolist()):
        assert c0[int(k)] == c
    assert np.allclose(re.points, mesh.points[re.point_data["tag"]])


def test_flat_vtk_cell_array_and_reference_mesh():
    mesh = embedded_tetra_mesh()
    assert mesh.cells.tolist() == [[0, 1, 2, 4], [0, 1, 4, 3], [0, 4, 2, 3], [4, 1, 2, 3]]
    assert np.allclose(_volumes(mesh), 1.0 / 24.0)


def test_attr_names_match_reference():
    assert MU.vtk == "mu" and LAMBDA.vtk == "lambda" and str(LAMBDA) == "lambda_"
    assert FIXED_MASK.vtk == "FixedMask" and GLOBAL_POINT_ID.vtk == "GlobalPointId"
    assert str(ACTIVATION) == "activation"
    la, mu = lame_converter(1.0e4, 0.3)
    assert abs(la - 1.0e4 * 0.3 / (1.3 * 0.4)) < 1e-9 and abs(mu - 1.0e4 / 2.6) < 1e-9


def test_dof_map_builder_and_maps():
    mesh = TetMesh(np.random.default_rng(0).random((6, 3)), np.array([[0, 1, 2, 3], [2, 3, 4, 5]]))
    b = DofMapBuilder()
    b.add_vertices(mesh)
    assert mesh.point_data["GlobalPointId"].tolist() == list(range(6))
    mask = np.zeros((6, 3), bool); mask[0] = True; mask[4, 1] = True
    val = np.zeros((6, 3)); val[0] = [1, 2, 3]; val[4, 1] = -5
    mesh.point_data["FixedMask"] = mask; mesh.point_data["FixedValue"] = val
    b.add_fixed(mesh)
    dm = b.finalize(dtype=torch.float64, device="cpu")
    assert isinstance(dm, DofMap) and dm.n_free == 14 and dm.n_fixed == 4 and dm.n_full == 18
    free = torch.arange(14, dtype=torch.float64)
    full = dm.to_full(free)
    assert full[0].tolist() == [1, 2, 3] and full[4, 1] == -5
    assert torch.equal(dm.to_free(full), free)
    assert (dm.to_full_grad(free)[torch.as_tensor(mask)] == 0).all()
    assert torch.equal(dm.free_mask(), ~torch.as_tensor(mask))
    # second mesh appended: ids continue (forward/dof_map/_builder.py:41-50)
    mesh2 = TetMesh(np.random.default_rng(1).random((4, 3)), np.array([[0, 1, 2, 3]]))
    b.add_vertices(mesh2)
    assert mesh2.point_data["GlobalPointId"].tolist() == [6, 7, 8, 9] and b.n_points == 10


def test_termination_classes():
    c = ConvergenceCriteria()
    assert _classify(1.0, 1e-9, c) is Result.PRIMARY_SUCCESS
    assert _classify(2.0, 1e-4, c) is Result.SECONDARY_SUCCESS
    assert _classify(2.0, 1e-1, c) is Result.MAX_STEPS_REACHED
    assert _classify(3.0, 1e-1, c) is Result.STAGNATION
    assert _classify(4.0, 1e-9, c) is Result.NAN_ENCOUNTERED


def test_region_matches_oracle_restatement():
    """apple_b200.fem.Region (closed form) == the oracle's literal restatement of Region.compute_grad
    (jax/fem/region/_region.py:84-108), including the reference's (cells, quadrature, ...) shapes."""
    from apple_b200.fem import Region
    from oracle import region as oregion

    mesh = cube_tet_mesh(5, grading=1.07)
    r = Region.from_pyvista(mesh, grad=True)
    dhdX, dV = oregion.compute_grad(mesh.points, mesh.cells)
    assert r.dhdX.shape == (mesh.n_cells, 1, 4, 3) and r.dV.shape == (mesh.n_cells, 1)
    np.testing.assert_allclose(r.dhdX[:, 0], dhdX, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(r.dV[:, 0], dV, rtol=1e-13)
    np.testing.assert_allclose(np.einsum("cqij,cqjk->cqik", r.dXdr, r.drdX), np.broadcast_to(np.eye(3), (mesh.n_cells, 1, 3, 3)), atol=1e-12)


def test_pcg_solves_the_hessian_system_of_the_oracle():
    """Jacobi-PCG on hess_prod (the adjoint solve) against a dense solve of the oracle's Hessian."""
    from apple_b200.optim import pcg
    from helpers import make_case, oracle_potential
    from oracle import fem as ofem, region as oregion

    mesh, _, _ = make_case(n=3, seed=2)
    V = mesh.n_points
    # rest state with lambda > mu / 3 in every cell: the (otherwise indefinite) SNH Hessian is SPD
    mesh.cell_data["lambda"] = 4.0 * mesh.cell_data["mu"]
    u = np.zeros((V, 3))
    model = ofem.Model([oracle_potential("snh", mesh)], V)
    fixed = np.zeros((V, 3), bool); fixed[mesh.points[:, 2] == 0.0] = True
    dm = oregion.DofMap(fixed, np.zeros((V, 3)))
    n = dm.n_free

    def matvec_np(v):
        return dm.to_free(model.hess_prod(u, dm.to_full_grad(v)))

    H = np.stack([matvec_np(e) for e in np.eye(n)], axis=1)
    assert np.linalg.eigvalsh(0.5 * (H + H.T)).min() > 0          # SPD near the rest state
    rng = np.random.default_rng(0)
    b = rng.standard_normal(n)
    M_inv = torch.from_numpy(1.0 / dm.to_free(model.hess_diag(u)))
    x, info = pcg(lambda v: torch.from_numpy(matvec_np(v.numpy())), torch.from_numpy(b), M_inv, tol=1e-10,
                  maxiter=4 * n, check_every=5)
    assert info.converged and info.residual_norm <= 1e-10 * info.rhs_norm
    np.testing.assert_allclose(x.numpy(), np.linalg.solve(H, b), rtol=1e-7, atol=1e-9)
    # unpreconditioned CG needs more iterations on this heterogeneous mesh
    _, plain = pcg(lambda v: torch.from_numpy(matvec_np(v.numpy())), torch.from_numpy(b), None, tol=1e-10,
                   maxiter=8 * n, check_every=5)
    assert plain.n_iters >= info.n_iters


def test_zeros_block_keeps_every_field_16_byte_aligned():
    """Outputs of a fused evaluation share one allocation (one memset); the kernels' vector REDs need every
    nodal field to start at a multiple of 16 bytes."""
    import torch

    from apple_b200.warp.model._adapter import zeros_block

    for dtype in (torch.float32, torch.float64):
        for n in (1, 5, 205_379):
            buf, fields, scalars = zeros_block(n, 3, 2, dtype, "cpu")
            assert len(fields) == 3 and len(scalars) == 2
            for f in fields:
                assert f.shape == (n, 3) and f.is_contiguous() and f.data_ptr() % 16 == 0
            ends = [f.data_ptr() + f.numel() * f.element_size() for f in fields]
            assert all(e <= nxt.data_ptr() for e, nxt in zip(ends, fields[1:] + [scalars[0]]))   # disjoint, ordered
            assert scalars[1].data_ptr() == scalars[0].data_ptr() + buf.element_size()
            fields[1].fill_(1.0); scalars[0].fill_(2.0)
            assert float(buf.sum()) == 3.0 * n + 2.0 and float(fields[0].sum()) == 0.0
            buf.zero_()
            assert float(fields[1].abs().sum()) == 0.0 and float(scalars[0]) == 0.0
