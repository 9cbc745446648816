"""``.vtu`` input / output without VTK (apple_b200/mesh/_vtu.py): round trips of the writer's three encodings and
hand-assembled files in the OTHER encodings vtkXMLUnstructuredGridWriter (pyvista's ``save``) can produce -- appended raw /
appended base64 data, zlib-compressed blocks, UInt32 headers -- with the reference's attribute names
(/root/reference/src/liblaf/apple/common/attr_name.py:36-44)."""

import base64
import struct
import zlib

import numpy as np
import pytest

from apple_b200.common import ACTIVATION, FIXED_MASK, FIXED_VALUE, FRACTION, LAMBDA, MU
from apple_b200.mesh import TetMesh, cube_tet_mesh, read_vtu, write_vtu


def _mesh():
    mesh = cube_tet_mesh(3, grading=1.1)
    rng = np.random.default_rng(0)
    T, V = mesh.n_cells, mesh.n_points
    mesh.cell_data[MU.vtk] = rng.uniform(1, 2, T)
    mesh.cell_data[LAMBDA.vtk] = rng.uniform(1, 2, T).astype(np.float32)
    mesh.cell_data[FRACTION.vtk] = rng.uniform(0.1, 1, T)
    mesh.cell_data[str(ACTIVATION)] = rng.standard_normal((T, 6))
    mesh.point_data[FIXED_MASK.vtk] = rng.random((V, 3)) < 0.2
    mesh.point_data[FIXED_VALUE.vtk] = rng.standard_normal((V, 3))
    mesh.point_data["GlobalPointId"] = np.arange(V, dtype=np.int32)
    return mesh


def _same(a: TetMesh, b: TetMesh, exact=True):
    assert a.n_points == b.n_points and a.n_cells == b.n_cells
    np.testing.assert_array_equal(a.cells, b.cells)
    cmp = np.testing.assert_array_equal if exact else (lambda x, y: np.testing.assert_allclose(x, y, rtol=1e-15))
    cmp(a.points, b.points)
    for da, db in ((a.point_data, b.point_data), (a.cell_data, b.cell_data)):
        assert set(da) == set(db)
        for k in da:
            x, y = np.asarray(da[k]), np.asarray(db[k])
            if x.dtype == bool:
                x = x.astype(np.uint8)
            cmp(x.reshape(y.shape), y)


@pytest.mark.parametrize("kw", [dict(binary=True), dict(binary=True, compress=True), dict(binary=False)],
                         ids=["base64", "base64+zlib", "ascii"])
def test_round_trip(tmp_path, kw):
    mesh = _mesh()
    path = tmp_path / "m.vtu"
    mesh.save(path, **kw)
    back = TetMesh.load(path)
    _same(mesh, back, exact=True)   # repr() of a double round-trips exactly, so even the ASCII file is bit-exact
    assert back.cell_data[LAMBDA.vtk].dtype == np.float32 and back.point_data["GlobalPointId"].dtype == np.int32


def _appended_file(mesh, path, raw: bool, compressed: bool, header: str):
    """The layout vtkXMLWriter uses in appended mode: DataArray elements carry offsets into one <AppendedData> blob."""
    hfmt = {"UInt32": "<I", "UInt64": "<Q"}[header]
    arrays, blob = [], b""

    def block(a):
        data = np.ascontiguousarray(a).tobytes()
        if not compressed:
            return struct.pack(hfmt, len(data)), data
        size = 1000
        chunks = [zlib.compress(data[i:i + size]) for i in range(0, len(data), size)]
        head = b"".join(struct.pack(hfmt, v) for v in (len(chunks), size, len(data) % size, *[len(c) for c in chunks]))
        return head, b"".join(chunks)

    def add(section, name, a, ncomp=None):
        nonlocal blob
        a = np.asarray(a)
        if a.dtype == bool:
            a = a.astype(np.uint8)
        vt = {"f8": "Float64", "f4": "Float32", "i8": "Int64", "i4": "Int32", "u1": "UInt8"}[a.dtype.str[1:]]
        head, body = block(a)
        # base64: one stream for an uncompressed block, header and blocks encoded separately for a compressed one
        enc = head + body if raw else (base64.b64encode(head) + base64.b64encode(body) if compressed else base64.b64encode(head + body))
        nc = f' NumberOfComponents="{ncomp}"' if ncomp else ""
        arrays.append((section, f'<DataArray type="{vt}" Name="{name}"{nc} format="appended" offset="{len(blob)}"/>'))
        blob += enc

    for k, v in mesh.point_data.items():
        v = np.asarray(v); add("PointData", k, v, v.shape[1] if v.ndim > 1 else None)
    for k, v in mesh.cell_data.items():
        v = np.asarray(v); add("CellData", k, v, v.shape[1] if v.ndim > 1 else None)
    add("Points", "Points", mesh.points, 3)
    T = mesh.n_cells
    add("Cells", "connectivity", mesh.cells.astype(np.int64).reshape(-1))
    add("Cells", "offsets", 4 * np.arange(1, T + 1, dtype=np.int64))
    add("Cells", "types", np.full(T, 10, np.uint8))
    comp = ' compressor="vtkZLibDataCompressor"' if compressed else ""
    xml = [f'<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="{header}"{comp}>',
           f'<UnstructuredGrid><Piece NumberOfPoints="{mesh.n_points}" NumberOfCells="{T}">']
    for sec in ("PointData", "CellData", "Points", "Cells"):
        xml.append(f"<{sec}>" + "".join(t for s, t in arrays if s == sec) + f"</{sec}>")
    xml.append("</Piece></UnstructuredGrid>")
    head = "\n".join(xml).encode() + (b'\n<AppendedData encoding="raw">\n_' if raw else b'\n<AppendedData encoding="base64">\n_')
    path.write_bytes(head + blob + b"\n</AppendedData>\n</VTKFile>\n")


@pytest.mark.parametrize("raw", [True, False], ids=["raw", "base64"])
@pytest.mark.parametrize("compressed", [False, True], ids=["plain", "zlib"])
@pytest.mark.parametrize("header", ["UInt32", "UInt64"])
def test_reads_the_appended_encodings_of_the_vtk_writer(tmp_path, raw, compressed, header):
    mesh = _mesh()
    path = tmp_path / "a.vtu"
    _appended_file(mesh, path, raw, compressed, header)
    _same(mesh, read_vtu(path))


def test_rejects_non_tetrahedral_files(tmp_path):
    mesh = _mesh()
    path = tmp_path / "m.vtu"
    write_vtu(mesh, path, binary=False)
    text = path.read_text().replace('Name="types" format="ascii">10 ', 'Name="types" format="ascii">12 ', 1)
    path.write_text(text)
    with pytest.raises(ValueError, match="tetrahedra"):
        read_vtu(path)


def test_loaded_mesh_feeds_the_region_exactly_like_the_generated_one(tmp_path):
    """from_pyvista on a mesh that went through a .vtu file: same rest shape and materials as the original."""
    from apple_b200.fem import Region

    mesh = _mesh()
    path = tmp_path / "m.vtu"
    mesh.save(path, compress=True)
    a, b = Region.from_pyvista(mesh, grad=True), Region.from_pyvista(TetMesh.load(path), grad=True)
    np.testing.assert_array_equal(a.dhdX, b.dhdX)
    np.testing.assert_array_equal(a.dV, b.dV)
    np.testing.assert_array_equal(a.cell_data[MU.vtk], b.cell_data[MU.vtk])
