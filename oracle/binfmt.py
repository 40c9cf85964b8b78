"""Oracle-side reader for the TorchSharp ``.bin`` checkpoint format (test infrastructure).

Follows Utils/Lib.cs:9-54 of /root/reference/YoloSharp: LEB128 tensor count, then per
tensor a .NET BinaryWriter string (7-bit-encoded length + UTF-8), LEB128 torch
ScalarType, LEB128 ndim, LEB128 dims, raw little-endian payload.
"""
import numpy as np
import torch

_DTYPES = {5: (np.float16, 2), 6: (np.float32, 4), 7: (np.float64, 8), 4: (np.int64, 8), 3: (np.int32, 4),
           15: (None, 2)}  # 15 = bfloat16 (no numpy dtype)


def _leb(buf, pos):
    num, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        num += (b & 0x7F) << shift
        if not b & 0x80:
            return num, pos
        shift += 7


def load_bin(path):
    """-> (ordered dict name -> torch tensor in file dtype, trailing byte count)."""
    buf = memoryview(open(path, "rb").read())
    pos = 0
    count, pos = _leb(buf, pos)
    out = {}
    for _ in range(count):
        ln, pos = _leb(buf, pos)
        name = bytes(buf[pos:pos + ln]).decode("utf-8")
        pos += ln
        dt, pos = _leb(buf, pos)
        nd, pos = _leb(buf, pos)
        shape = []
        for _ in range(nd):
            d, pos = _leb(buf, pos)
            shape.append(d)
        npdt, isz = _DTYPES[dt]
        n = int(np.prod(shape)) if shape else 1
        raw = bytes(buf[pos:pos + n * isz])
        pos += n * isz
        if dt == 15:
            t = torch.frombuffer(bytearray(raw), dtype=torch.bfloat16).reshape(shape)
        else:
            t = torch.from_numpy(np.frombuffer(raw, dtype=npdt).copy().reshape(shape))
        out[name] = t
    return out, len(buf) - pos


def load_into(model, path):
    """Load a .bin into an oracle model (fp32).  Returns (missing, unexpected)."""
    sd, trailing = load_bin(path)
    assert trailing == 0, f"{trailing} trailing bytes"
    own = model.state_dict()
    missing = [k for k in own if k not in sd]
    unexpected = [k for k in sd if k not in own]
    new = {}
    for k, v in sd.items():
        if k in own:
            if own[k].shape != v.shape:
                if own[k].numel() == v.numel():
                    v = v.reshape(own[k].shape)  # num_batches_tracked etc.
                else:
                    raise ValueError(f"shape mismatch {k}: {tuple(v.shape)} vs {tuple(own[k].shape)}")
            new[k] = v.to(own[k].dtype)
    model.load_state_dict(new, strict=False)
    return missing, unexpected
