"""VTK ImageData (.vti) loader without the ``vtk`` package.

Drop-in for ``pyvr.dataloaders.load_vtk_volume`` (reference ``pyvr/dataloaders/vtk_loader.py:15-146``):
same signature, name format, min-max normalisation, bounds rule (longest physical side -> [-1, 1],
centred at the origin), automatic normals and error types.  The reference reads the file through
``vtkXMLImageDataReader`` (a C++ wheel that is absent from this image); this module decodes the XML
container itself with the standard library: ``format`` = appended / binary / ascii, ``encoding`` =
base64 / raw, ``compressor`` = none / vtkZLibDataCompressor, ``header_type`` = UInt32 / UInt64.

Quirk kept on purpose: the array is shaped ``(nz, ny, nx)`` (reference ``vtk_loader.py:102-103``) while
the bounds are computed in VTK (x, y, z) order (``:117-129``), so VTK's z axis is rendered along world x.
That is self-consistent only for cubic data, which both reference fixtures are.

Normals are computed by the sm_100a stencil (``Volume.compute_normals``), so the default call needs a
GPU; pass ``compute_normals=False`` on a CPU-only box.
"""

from __future__ import annotations

import base64
import re
import struct
import zlib
from pathlib import Path
from typing import Union
from xml.etree import ElementTree

import numpy as np

from .volume import Volume

_VTK_DTYPES = {
    "Int8": "i1", "UInt8": "u1", "Int16": "i2", "UInt16": "u2", "Int32": "i4", "UInt32": "u4",
    "Int64": "i8", "UInt64": "u8", "Float32": "f4", "Float64": "f8",
}


def _b64_len(nbytes: int) -> int:
    return (nbytes + 2) // 3 * 4


def _decode_block_stream(text: bytes, offset_chars: int, compressed: bool, hdr_fmt: str) -> bytes:
    """Decode one base64 DataArray payload starting ``offset_chars`` into ``text``."""
    hsize = struct.calcsize(hdr_fmt)
    if not compressed:
        # [nbytes][data], header and data base64-encoded as one stream (VTK >= 4.x) or separately
        head = base64.b64decode(text[offset_chars:offset_chars + _b64_len(hsize)])
        (nbytes,) = struct.unpack(hdr_fmt, head[:hsize])
        joint = text[offset_chars:offset_chars + _b64_len(hsize + nbytes)]
        data = base64.b64decode(joint)[hsize:hsize + nbytes]
        if len(data) == nbytes:
            return data
        start = offset_chars + _b64_len(hsize)
        return base64.b64decode(text[start:start + _b64_len(nbytes)])[:nbytes]
    # compressed: header [nblocks, blocksize, last_blocksize, csize_0 ... csize_{n-1}] is its own base64
    # stream, followed by the concatenated zlib blocks as a second base64 stream
    first = base64.b64decode(text[offset_chars:offset_chars + _b64_len(3 * hsize)])
    nblocks, blocksize, last = struct.unpack("<3" + hdr_fmt[-1], first[:3 * hsize])
    head_chars = _b64_len((3 + nblocks) * hsize)
    head = base64.b64decode(text[offset_chars:offset_chars + head_chars])
    csizes = struct.unpack(f"<{nblocks}" + hdr_fmt[-1], head[3 * hsize:(3 + nblocks) * hsize])
    start = offset_chars + head_chars
    blob = base64.b64decode(text[start:start + _b64_len(sum(csizes))])
    out, pos = [], 0
    for size in csizes:
        out.append(zlib.decompress(blob[pos:pos + size]))
        pos += size
    data = b"".join(out)
    expect = (nblocks - 1) * blocksize + (last if last else blocksize) if nblocks else 0
    if len(data) != expect:
        raise ValueError(f"corrupt compressed block stream: {len(data)} bytes, header says {expect}")
    return data


def _decode_raw_appended(raw: bytes, offset: int, compressed: bool, hdr_fmt: str) -> bytes:
    hsize = struct.calcsize(hdr_fmt)
    if not compressed:
        (nbytes,) = struct.unpack(hdr_fmt, raw[offset:offset + hsize])
        return raw[offset + hsize:offset + hsize + nbytes]
    nblocks, blocksize, last = struct.unpack("<3" + hdr_fmt[-1], raw[offset:offset + 3 * hsize])
    csizes = struct.unpack(f"<{nblocks}" + hdr_fmt[-1], raw[offset + 3 * hsize:offset + (3 + nblocks) * hsize])
    pos = offset + (3 + nblocks) * hsize
    out = []
    for size in csizes:
        out.append(zlib.decompress(raw[pos:pos + size]))
        pos += size
    return b"".join(out)


def read_vti(file_path: Union[str, Path]):
    """Parse a .vti file.  Returns ``(dims_xyz, spacing_xyz, {array_name: (flat ndarray, n_components)})``
    for the PointData arrays of the first piece."""
    raw = Path(file_path).read_bytes()
    # Raw appended data is not valid XML: cut it out before parsing.
    appended_raw = None
    m = re.search(rb"<AppendedData[^>]*encoding=\"raw\"[^>]*>", raw)
    if m:
        start = raw.index(b"_", m.end()) + 1
        end = raw.rindex(b"</AppendedData>")
        appended_raw = raw[start:end]
        raw = raw[:m.end()] + b"_" + raw[end:]
    try:
        root = ElementTree.fromstring(raw)
    except ElementTree.ParseError as exc:
        raise ValueError(f"Not a valid VTK XML file: {file_path}: {exc}") from exc
    if root.tag != "VTKFile" or root.get("type") != "ImageData":
        raise ValueError(f"Not a VTK ImageData file: {file_path}")
    if root.get("byte_order", "LittleEndian") != "LittleEndian":
        raise ValueError("Only little-endian .vti files are supported")
    hdr_fmt = "<Q" if root.get("header_type", "UInt32") == "UInt64" else "<I"
    compressor = root.get("compressor")
    if compressor not in (None, "", "vtkZLibDataCompressor"):
        raise ValueError(f"Unsupported compressor: {compressor}")
    compressed = bool(compressor)

    image = root.find("ImageData")
    if image is None:
        raise ValueError(f"No ImageData element in {file_path}")
    ext = [int(v) for v in image.get("WholeExtent").split()]
    dims = (ext[1] - ext[0] + 1, ext[3] - ext[2] + 1, ext[5] - ext[4] + 1)
    spacing = tuple(float(v) for v in image.get("Spacing", "1 1 1").split())

    appended = root.find("AppendedData")
    appended_text = None
    if appended is not None and appended_raw is None:
        text = (appended.text or "").strip()
        appended_text = text[text.index("_") + 1:].encode() if "_" in text else b""

    arrays = {}
    piece = image.find("Piece")
    point_data = piece.find("PointData") if piece is not None else None
    for arr in (point_data.findall("DataArray") if point_data is not None else []):
        dtype = _VTK_DTYPES.get(arr.get("type"))
        if dtype is None:
            raise ValueError(f"Unsupported DataArray type: {arr.get('type')}")
        ncomp = int(arr.get("NumberOfComponents", "1"))
        fmt = arr.get("format", "ascii")
        if fmt == "ascii":
            flat = np.array((arr.text or "").split(), dtype="<" + dtype)
        elif fmt == "binary":
            payload = "".join((arr.text or "").split()).encode()
            flat = np.frombuffer(_decode_block_stream(payload, 0, compressed, hdr_fmt), dtype="<" + dtype)
        elif fmt == "appended":
            off = int(arr.get("offset", "0"))
            if appended_raw is not None:
                data = _decode_raw_appended(appended_raw, off, compressed, hdr_fmt)
            else:
                data = _decode_block_stream(appended_text, off, compressed, hdr_fmt)
            flat = np.frombuffer(data, dtype="<" + dtype)
        else:
            raise ValueError(f"Unsupported DataArray format: {fmt}")
        arrays[arr.get("Name")] = (flat, ncomp)
    return dims, spacing, arrays


def load_vtk_volume(file_path: Union[str, Path], scalars_name: str = "Scalars_",
                    compute_normals: bool = True) -> Volume:
    """Load a VTK ImageData (.vti) file as a :class:`Volume` (see the module docstring)."""
    file_path = Path(file_path)
    if not file_path.exists():
        raise FileNotFoundError(f"VTK file not found: {file_path}")
    dims, spacing, arrays = read_vti(file_path)
    if any(d <= 0 for d in dims):
        raise ValueError(f"Invalid dimensions: {dims}")
    if any(s <= 0 for s in spacing):
        raise ValueError(f"Invalid spacing (must be positive): {spacing}")
    if scalars_name not in arrays:
        raise ValueError(f"Scalar array '{scalars_name}' not found in {file_path}. "
                         f"Available arrays: {list(arrays)}")
    flat, ncomp = arrays[scalars_name]
    if ncomp != 1:
        raise ValueError(f"Multi-component scalars not supported. "
                         f"Array '{scalars_name}' has {ncomp} components, expected 1.")
    if flat.size != dims[0] * dims[1] * dims[2]:
        raise ValueError(f"Array '{scalars_name}' holds {flat.size} values, extent needs {dims[0] * dims[1] * dims[2]}")

    data = flat.reshape((dims[2], dims[1], dims[0])).astype(np.float32)      # (nz, ny, nx)
    lo, hi = data.min(), data.max()
    normalized = np.zeros_like(data) if hi - lo < 1e-9 else (data - lo) / (hi - lo)

    physical = np.array([dims[0] * spacing[0], dims[1] * spacing[1], dims[2] * spacing[2]], dtype=np.float32)
    half = physical * (2.0 / np.max(physical)) / 2.0
    volume = Volume(data=normalized, normals=None, min_bounds=-half, max_bounds=+half,
                    name=f"{file_path.name}({scalars_name})")
    if compute_normals:
        volume.compute_normals()
    return volume
