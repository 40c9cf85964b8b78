"""Checkpoint conversion with the library's native readers / writer (csrc/ckpt.cu; no GPU needed): torch.save `.pt` / `.pth`,
`.safetensors` or TorchSharp `.bin`  ->  TorchSharp `.bin`, the format `YoloTask.LoadModel` reads (Utils/Lib.cs:9-54).
The reference does this with its TorchSharp model in the loop (Tools.TransModelFromSafetensors, Utils/Tools.cs:16-35: load into the
model, `model.save`); here the tensors are written as the source names them.
  python tools/convert_checkpoint.py src.pt dst.bin [--strip model.] [--half]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def convert(src, dst, strip="", half=False):
    import torch

    from yolosharp_b200 import engine as E
    sd = E.read_checkpoint(src)
    out = {}
    for k, v in sd.items():
        if strip and k.startswith(strip):
            k = k[len(strip):]
        if half and v.dtype in (torch.float32, torch.bfloat16, torch.float64):
            v = v.to(torch.float16)
        out[k] = v
    E.write_checkpoint_bin(dst, out)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--strip", default="", help="drop this prefix from tensor names (an Ultralytics checkpoint: 'model.')")
    ap.add_argument("--half", action="store_true", help="store floating-point tensors as float16 (the shipped .bin files are)")
    a = ap.parse_args()
    o = convert(a.src, a.dst, a.strip, a.half)
    print(f"{a.dst}: {len(o)} tensors")
