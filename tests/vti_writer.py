"""Test helper: write VTK XML ImageData (.vti) files in every container variant the std-lib reader
(pyvr_b200/dataloaders.py) understands.  Independent of the reader: it builds the byte streams from
the VTK file-format description (header [nblocks, blocksize, lastsize, csizes...] + zlib blocks, or
[nbytes] + data), so reader and writer do not share code."""

import base64
import struct
import zlib

import numpy as np

_VTK_TYPE = {"float32": "Float32", "float64": "Float64", "uint8": "UInt8", "int16": "Int16", "uint16": "UInt16",
             "int32": "Int32"}


def _payload(data: bytes, compressed: bool, hdr: str, block: int = 32768):
    """(header bytes, body bytes) of one DataArray."""
    if not compressed:
        return struct.pack(hdr, len(data)), data
    blocks = [data[i:i + block] for i in range(0, len(data), block)] or [b""]
    comp = [zlib.compress(b) for b in blocks]
    last = len(blocks[-1]) if len(blocks[-1]) != block else 0
    head = struct.pack("<3" + hdr[-1], len(blocks), block, last) + b"".join(struct.pack(hdr, len(c)) for c in comp)
    return head, b"".join(comp)


def write_vti(path, arrays, dims_xyz, spacing=(1.0, 1.0, 1.0), fmt="appended", encoding="base64",
              compressed=True, header_type="UInt32", components=None):
    """arrays: {name: flat ndarray in VTK order (x fastest)}; fmt: appended | binary | ascii."""
    hdr = "<Q" if header_type == "UInt64" else "<I"
    nx, ny, nz = dims_xyz
    attrs = f'type="ImageData" version="1.0" byte_order="LittleEndian" header_type="{header_type}"'
    if compressed and fmt != "ascii":
        attrs += ' compressor="vtkZLibDataCompressor"'
    out = [f'<?xml version="1.0"?>\n<VTKFile {attrs}>\n'.encode(),
           f'  <ImageData WholeExtent="0 {nx - 1} 0 {ny - 1} 0 {nz - 1}" Origin="0 0 0" '
           f'Spacing="{spacing[0]} {spacing[1]} {spacing[2]}">\n'
           f'  <Piece Extent="0 {nx - 1} 0 {ny - 1} 0 {nz - 1}">\n    <PointData Scalars="{next(iter(arrays))}">\n'.encode()]
    appended = b""
    for name, arr in arrays.items():
        arr = np.ascontiguousarray(arr)
        ncomp = (components or {}).get(name, 1)
        common = f'type="{_VTK_TYPE[arr.dtype.name]}" Name="{name}" NumberOfComponents="{ncomp}"'
        if fmt == "ascii":
            text = " ".join(repr(float(v)) if arr.dtype.kind == "f" else str(int(v)) for v in arr.ravel())
            out.append(f'      <DataArray {common} format="ascii">\n{text}\n      </DataArray>\n'.encode())
            continue
        head, body = _payload(arr.tobytes(), compressed, hdr)
        if fmt == "binary":
            enc = (base64.b64encode(head) + base64.b64encode(body)) if compressed else base64.b64encode(head + body)
            out.append(f'      <DataArray {common} format="binary">\n'.encode() + enc + b'\n      </DataArray>\n')
        else:
            if encoding == "raw":
                off, chunk = len(appended), head + body
            else:
                off = len(appended)
                chunk = (base64.b64encode(head) + base64.b64encode(body)) if compressed else base64.b64encode(head + body)
            out.append(f'      <DataArray {common} format="appended" offset="{off}"/>\n'.encode())
            appended += chunk
    out.append(b'    </PointData>\n    <CellData>\n    </CellData>\n  </Piece>\n  </ImageData>\n')
    if fmt == "appended":
        out.append(f'  <AppendedData encoding="{encoding}">\n   _'.encode() + appended + b'\n  </AppendedData>\n')
    out.append(b'</VTKFile>\n')
    with open(path, "wb") as f:
        f.write(b"".join(out))
    return path
