"""VTK XML UnstructuredGrid (``.vtu``) input / output for tetrahedral meshes, without VTK or pyvista.

The reference reads and writes its meshes through pyvista (``melon.load_unstructured_grid`` /
``mesh.save`` in ``exp/**``; attribute names ``common/attr_name.py:36-44``); neither pyvista nor VTK is installable
here, so this is a small stand-alone implementation of the file format (VTK file formats, "XML formats"):

* reader: ``<DataArray format="ascii" | "binary" | "appended">``, base64 and raw appended data, optional
  ``vtkZLibDataCompressor``, ``header_type`` UInt32 / UInt64, little- and big-endian files -- i.e. what
  ``vtkXMLUnstructuredGridWriter`` (pyvista's ``save``) produces in any of its modes; cells that are not linear
  tetrahedra (``types != 10``) are rejected;
* writer: inline base64 (``binary=True``, optionally zlib-compressed) or ASCII.

Data arrays keep their names, so ``cell_data["mu" | "lambda" | "Fraction" | "activation"]`` and
``point_data["FixedMask" | "FixedValue" | "GlobalPointId" | "Force"]`` round-trip unchanged.
"""

from __future__ import annotations

import base64
import struct
import xml.etree.ElementTree as ET
import zlib
from pathlib import Path

import numpy as np

from ._mesh import VTK_TETRA, TetMesh

_VTK_TO_NP = {
    "Int8": "i1", "UInt8": "u1", "Int16": "i2", "UInt16": "u2", "Int32": "i4", "UInt32": "u4", "Int64": "i8", "UInt64": "u8",
    "Float32": "f4", "Float64": "f8",
}
_NP_TO_VTK = {np.dtype(v).str[1:]: k for k, v in _VTK_TO_NP.items()}


def _decode_blocks(raw: bytes, header_dtype: np.dtype, compressed: bool) -> bytes:
    """One data block of the XML formats: ``[header][payload]`` (see the VTK file-formats document)."""
    hs = header_dtype.itemsize
    if not compressed:
        (n,) = np.frombuffer(raw[:hs], header_dtype)
        return raw[hs:hs + int(n)]
    nblocks, _, _ = (int(v) for v in np.frombuffer(raw[:3 * hs], header_dtype))
    sizes = [int(v) for v in np.frombuffer(raw[3 * hs:(3 + nblocks) * hs], header_dtype)]
    out, pos = [], (3 + nblocks) * hs
    for s in sizes:
        out.append(zlib.decompress(raw[pos:pos + s]))
        pos += s
    return b"".join(out)


def _b64_block(text: str, header_dtype: np.dtype, compressed: bool) -> bytes:
    """Inline / appended base64.  Uncompressed: the length prefix and the payload are ONE base64 stream; compressed: the
    block header is encoded on its own (padded), the compressed blocks follow as a second stream (the convention of
    vtkXMLWriter, as also implemented by meshio)."""
    text = "".join(text.split())
    hs = header_dtype.itemsize
    if not compressed:
        (n,) = np.frombuffer(base64.b64decode(text[: -(-hs // 3) * 4])[:hs], header_dtype)
        total = -(-(hs + int(n)) // 3) * 4
        return base64.b64decode(text[:total])[hs:hs + int(n)]
    first = base64.b64decode(text[: -(-3 * hs // 3) * 4])
    nblocks = int(np.frombuffer(first[:hs], header_dtype)[0])
    hlen = (3 + nblocks) * hs
    hb64 = -(-hlen // 3) * 4
    header = base64.b64decode(text[:hb64])[:hlen]
    sizes = [int(v) for v in np.frombuffer(header[3 * hs:], header_dtype)]
    body = base64.b64decode(text[hb64:hb64 + -(-sum(sizes) // 3) * 4])
    out, pos = [], 0
    for s in sizes:
        out.append(zlib.decompress(body[pos:pos + s]))
        pos += s
    return b"".join(out)


class _Reader:
    def __init__(self, path):
        data = Path(path).read_bytes()
        self.appended_raw = None
        # raw appended data is not valid XML: cut it out before parsing
        marker = data.find(b"<AppendedData")
        if marker >= 0:
            tag_end = data.find(b">", marker) + 1
            enc = b'encoding="raw"' in data[marker:tag_end]
            close = data.rfind(b"</AppendedData>")
            blob = data[tag_end:close]
            under = blob.find(b"_")
            self.appended = blob[under + 1:]
            self.appended_is_raw = enc
            data = data[:tag_end] + b"</AppendedData>" + data[close + len(b"</AppendedData>"):]
        self.root = ET.fromstring(data)
        if self.root.attrib.get("type") != "UnstructuredGrid":
            raise ValueError("not a VTK XML UnstructuredGrid file")
        self.order = "<" if self.root.attrib.get("byte_order", "LittleEndian") == "LittleEndian" else ">"
        self.header_dtype = np.dtype(self.order + _VTK_TO_NP[self.root.attrib.get("header_type", "UInt32")])
        comp = self.root.attrib.get("compressor", "")
        if comp and comp != "vtkZLibDataCompressor":
            raise ValueError(f"unsupported compressor {comp}")
        self.compressed = bool(comp)

    def array(self, node) -> np.ndarray:
        dt = np.dtype(self.order + _VTK_TO_NP[node.attrib["type"]])
        ncomp = int(node.attrib.get("NumberOfComponents", "1"))
        fmt = node.attrib.get("format", "ascii")
        if fmt == "ascii":
            a = np.array(node.text.split(), dtype=np.float64 if dt.kind == "f" else np.int64).astype(dt.newbyteorder("="))
        elif fmt == "binary":
            a = np.frombuffer(_b64_block(node.text, self.header_dtype, self.compressed), dt)
        elif fmt == "appended":
            off = int(node.attrib["offset"])
            if self.appended_is_raw:
                hs = self.header_dtype.itemsize
                raw = self.appended[off:]
                if self.compressed:
                    nblocks = int(np.frombuffer(raw[:hs], self.header_dtype)[0])
                    sizes = np.frombuffer(raw[3 * hs:(3 + nblocks) * hs], self.header_dtype)
                    raw = raw[:(3 + nblocks) * hs + int(sizes.sum())]
                a = np.frombuffer(_decode_blocks(raw, self.header_dtype, self.compressed), dt)
            else:
                a = np.frombuffer(_b64_block(self.appended[off:].decode("ascii"), self.header_dtype, self.compressed), dt)
        else:
            raise ValueError(f"unknown DataArray format {fmt}")
        a = a.astype(dt.newbyteorder("="))
        return a.reshape(-1, ncomp) if ncomp > 1 else a


def read_vtu(path) -> TetMesh:
    """A ``.vtu`` file of linear tetrahedra -> ``TetMesh`` with every point / cell data array."""
    r = _Reader(path)
    piece = r.root.find("UnstructuredGrid/Piece")
    if piece is None:
        raise ValueError("no <Piece> in the file")
    n_points, n_cells = int(piece.attrib["NumberOfPoints"]), int(piece.attrib["NumberOfCells"])
    points = r.array(piece.find("Points/DataArray")).reshape(n_points, 3)
    cells = {n.attrib["Name"]: r.array(n) for n in piece.findall("Cells/DataArray")}
    types = cells["types"]
    if n_cells and not np.all(types == VTK_TETRA):
        raise ValueError("only linear tetrahedra (VTK cell type 10) are supported")
    offsets = cells["offsets"].astype(np.int64)
    if n_cells and not np.array_equal(offsets, 4 * np.arange(1, n_cells + 1)):
        raise ValueError("connectivity offsets do not describe 4-node cells")
    conn = cells["connectivity"].astype(np.int64).reshape(n_cells, 4)

    def data(tag):
        node = piece.find(tag)
        return {} if node is None else {n.attrib["Name"]: r.array(n) for n in node.findall("DataArray")}

    return TetMesh(points, conn, point_data=data("PointData"), cell_data=data("CellData"))


def _encode(a: np.ndarray, binary: bool, compress: bool) -> str:
    if not binary:
        flat = a.reshape(-1)
        return " ".join(repr(float(v)) if a.dtype.kind == "f" else str(int(v)) for v in flat)
    raw = np.ascontiguousarray(a).astype(a.dtype.newbyteorder("<")).tobytes()
    if not compress:
        return base64.b64encode(struct.pack("<Q", len(raw)) + raw).decode("ascii")
    block = 1 << 15
    chunks = [zlib.compress(raw[i:i + block]) for i in range(0, len(raw), block)] or [zlib.compress(b"")]
    last = len(raw) % block
    header = struct.pack(f"<{3 + len(chunks)}Q", len(chunks), block, last, *[len(c) for c in chunks])
    return (base64.b64encode(header) + base64.b64encode(b"".join(chunks))).decode("ascii")


def write_vtu(mesh, path, *, binary: bool = True, compress: bool = False) -> None:
    """Writes a ``TetMesh`` (or anything with ``points``, ``cells`` (n, 4), ``point_data``, ``cell_data``) as ``.vtu``."""
    points = np.ascontiguousarray(mesh.points, dtype=np.float64)
    cells = np.ascontiguousarray(mesh.cells, dtype=np.int64).reshape(-1, 4)
    n_points, n_cells = points.shape[0], cells.shape[0]
    root = ET.Element("VTKFile", type="UnstructuredGrid", version="1.0", byte_order="LittleEndian", header_type="UInt64")
    if binary and compress:
        root.set("compressor", "vtkZLibDataCompressor")
    piece = ET.SubElement(ET.SubElement(root, "UnstructuredGrid"), "Piece", NumberOfPoints=str(n_points),
                          NumberOfCells=str(n_cells))

    def add(parent, name, a):
        a = np.asarray(a)
        if a.dtype == bool:
            a = a.astype(np.uint8)
        key = a.dtype.str[1:]
        if key not in _NP_TO_VTK:
            raise TypeError(f"data array {name!r}: unsupported dtype {a.dtype}")
        node = ET.SubElement(parent, "DataArray", type=_NP_TO_VTK[key], Name=name, format="binary" if binary else "ascii")
        if a.ndim > 1:
            node.set("NumberOfComponents", str(int(np.prod(a.shape[1:]))))
        node.text = _encode(a, binary, compress)

    pd, cd = ET.SubElement(piece, "PointData"), ET.SubElement(piece, "CellData")
    for k, v in mesh.point_data.items():
        add(pd, k, np.asarray(v).reshape(n_points, -1) if np.asarray(v).ndim > 1 else v)
    for k, v in mesh.cell_data.items():
        add(cd, k, np.asarray(v).reshape(n_cells, -1) if np.asarray(v).ndim > 1 else v)
    add(ET.SubElement(piece, "Points"), "Points", points)
    cn = ET.SubElement(piece, "Cells")
    add(cn, "connectivity", cells.reshape(-1))
    add(cn, "offsets", 4 * np.arange(1, n_cells + 1, dtype=np.int64))
    add(cn, "types", np.full(n_cells, VTK_TETRA, np.uint8))
    ET.ElementTree(root).write(str(path), xml_declaration=True, encoding="utf-8")
