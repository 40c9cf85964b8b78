"""Native checkpoint reader / writer (csrc/ckpt.cu) against the independent Python reader (yolosharp_b200/binfmt.py),
the committed copy of the reference's shipped Yolov8n.bin (tests/golden/yolov8n_f16.npz) and, when the reference tree is
mounted, the shipped file itself (byte-exact round trip)."""
import json
import os
import struct

import numpy as np
import pytest
import torch

from tests.util import GOLDEN
from yolosharp_b200 import binfmt
from yolosharp_b200 import engine as E
from yolosharp_b200._lib import YbError

REF_BIN = "/root/reference/YoloSharpDemo/Assets/PreTrainedModels/Yolov8n.bin"
CODE = {torch.float16: 5, torch.float32: 6, torch.int64: 4, torch.int32: 3}


def golden_sd():
    z = np.load(os.path.join(GOLDEN, "yolov8n_f16.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_bin_native_reader_equals_python_reader(tmp_path):
    sd = golden_sd()
    p = str(tmp_path / "a.bin")
    binfmt.write_bin(p, [(k, CODE[v.dtype], list(v.shape), v.numpy().tobytes()) for k, v in sd.items()])
    got = E.read_checkpoint(p)
    assert list(got) == list(sd) and len(got) == 357
    for k, v in sd.items():
        assert got[k].dtype == v.dtype and tuple(got[k].shape) == tuple(v.shape) and torch.equal(got[k], v), k


def test_bin_native_writer_equals_python_writer_and_reference_file(tmp_path):
    sd = golden_sd()
    a, b = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    binfmt.write_bin(a, [(k, CODE[v.dtype], list(v.shape), v.numpy().tobytes()) for k, v in sd.items()])
    E.write_checkpoint_bin(b, sd)
    assert open(a, "rb").read() == open(b, "rb").read()
    if os.path.exists(REF_BIN):  # authoring container only: the shipped file round-trips byte for byte
        c = str(tmp_path / "c.bin")
        E.write_checkpoint_bin(c, E.read_checkpoint(REF_BIN))
        assert open(c, "rb").read() == open(REF_BIN, "rb").read()


def test_safetensors_reader(tmp_path):
    sd = {k: v for k, v in list(golden_sd().items())[:20] if v.numel()}
    sd["extra.bf16"] = torch.randn(3, 5).to(torch.bfloat16)
    hdr, blob = {"__metadata__": {"format": "pt"}}, b""
    names = {torch.float16: "F16", torch.float32: "F32", torch.int64: "I64", torch.bfloat16: "BF16"}
    for k, v in sd.items():
        raw = v.view(torch.int16).numpy().tobytes() if v.dtype == torch.bfloat16 else v.numpy().tobytes()
        hdr[k] = {"dtype": names[v.dtype], "shape": list(v.shape), "data_offsets": [len(blob), len(blob) + len(raw)]}
        blob += raw
    hj = json.dumps(hdr).encode()
    p = str(tmp_path / "m.safetensors")
    open(p, "wb").write(struct.pack("<Q", len(hj)) + hj + blob)
    got = E.read_checkpoint(p)
    assert set(got) == set(sd)
    for k, v in sd.items():
        assert got[k].dtype == v.dtype and torch.equal(got[k], v), k


def test_checkpoint_errors(tmp_path):
    with pytest.raises(YbError):
        E.read_checkpoint(str(tmp_path / "missing.bin"))
    bad = str(tmp_path / "bad.safetensors")
    open(bad, "wb").write(b"\x05\x00\x00\x00\x00\x00\x00\x00{\"a\"")
    with pytest.raises(YbError):
        E.read_checkpoint(bad)
    trunc = str(tmp_path / "t.bin")
    sd = golden_sd()
    full = str(tmp_path / "f.bin")
    E.write_checkpoint_bin(full, sd)
    open(trunc, "wb").write(open(full, "rb").read()[:-7])
    with pytest.raises(YbError):
        E.read_checkpoint(trunc)
    with pytest.raises(YbError) as ei:
        E.read_checkpoint(str(tmp_path / "model.pt"))
    assert "cannot read" in str(ei.value) or "pickle" in str(ei.value)


def test_engine_load_checkpoint_dry_run(tmp_path):
    """yb_load_checkpoint on a dry-run engine (no GPU): every expected tensor of Yolov8n is found in the file."""
    import yolosharp_b200 as y
    from yolosharp_b200 import _lib as L
    p = str(tmp_path / "n.bin")
    E.write_checkpoint_bin(p, golden_sd())
    e = y.Engine("v8", "n", "detect", 80, "f16", 0, 1, 64, 64, flags=L.YB_FLAG_DRY_RUN)
    loaded, missing = e.load_checkpoint(p)
    assert missing == 0 and loaded >= len(e.expected_tensors())
    e2 = y.Engine("v8", "s", "detect", 80, "f16", 0, 1, 64, 64, flags=L.YB_FLAG_DRY_RUN)
    with pytest.raises(YbError):  # n-size file into an s-size graph: shapes are checked at finalize; here names match, so
        e2.load_checkpoint(p)     # loading succeeds and ...
        e2.finalize()             # ... finalize refuses (dry-run engine / shape mismatch)
    e.close()
    e2.close()


def test_train_state_dict_roundtrip(tmp_path):
    """trained weights go back to the reference checkpoint format (SaveWeight) and load again."""
    from tests.torch_train_ops import TorchOps
    from tests.util import oracle_model
    from yolosharp_b200.train import TrainStepV8
    m = oracle_model("v8", "detect", "n")
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    st = TrainStepV8(sd0, "n", 80, device="cpu", ops=TorchOps(), lr=1e-3)
    st.step_count = 3
    p = str(tmp_path / "last.bin")
    st.save(p)
    back = E.read_checkpoint(p)
    assert set(back) == set(sd0), (set(sd0) ^ set(back))
    for k, v in sd0.items():
        if k.endswith("num_batches_tracked"):
            assert int(back[k]) == 3
        elif v.numel():
            assert torch.equal(back[k].float(), v.float()), k
