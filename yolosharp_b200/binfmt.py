"""Reader/writer for the reference's TorchSharp ``.bin`` checkpoint format (host side).

Format as consumed by Utils/Lib.cs:9-54 and produced by Models/YoloBaseTaskModel.cs:470-490 of the
reference: LEB128 count, then per tensor {.NET 7-bit-length-prefixed UTF-8 name, LEB128 torch
ScalarType, LEB128 ndim, LEB128 dims, raw little-endian data}.  Tensors are returned as
(dtype code, shape, bytes) so they can be handed to ``yb_load_tensor`` without a round trip through
any tensor library.
"""
import struct

ITEMSIZE = {0: 1, 1: 1, 2: 2, 3: 4, 4: 8, 5: 2, 6: 4, 7: 8, 15: 2}


def _read_leb(f):
    num = shift = 0
    while True:
        b = f.read(1)
        if not b:
            raise EOFError("truncated .bin file")
        num |= (b[0] & 0x7F) << shift
        if not b[0] & 0x80:
            return num
        shift += 7


def _write_leb(f, v):
    while True:
        b = v & 0x7F
        v >>= 7
        f.write(bytes([b | (0x80 if v else 0)]))
        if not v:
            return


def read_bin(path):
    """-> list of (name, dtype_code, shape tuple, payload bytes) in file order."""
    out = []
    with open(path, "rb") as f:
        count = _read_leb(f)
        for _ in range(count):
            name = f.read(_read_leb(f)).decode("utf-8")
            dt = _read_leb(f)
            shape = tuple(_read_leb(f) for _ in range(_read_leb(f)))
            if dt not in ITEMSIZE:
                raise ValueError(f"{path}: unsupported scalar type {dt} for {name}")
            n = 1
            for d in shape:
                n *= d
            data = f.read(n * ITEMSIZE[dt])
            if len(data) != n * ITEMSIZE[dt]:
                raise EOFError(f"{path}: truncated payload for {name}")
            out.append((name, dt, shape, data))
        if f.read(1):
            raise ValueError(f"{path}: trailing bytes after {count} tensors")
    return out


def write_bin(path, tensors):
    """tensors: iterable of (name, dtype_code, shape, payload bytes)."""
    tensors = list(tensors)
    with open(path, "wb") as f:
        _write_leb(f, len(tensors))
        for name, dt, shape, data in tensors:
            nb = name.encode("utf-8")
            _write_leb(f, len(nb))
            f.write(nb)
            _write_leb(f, dt)
            _write_leb(f, len(shape))
            for d in shape:
                _write_leb(f, d)
            f.write(data)
