"""Time of one training step on the fp32 parity path (yolosharp_b200/train.py): python tools/exp_train_time.py [B] [size]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests.test_train_step import _targets  # noqa: E402
from tests.util import oracle_model, synth_image  # noqa: E402
from yolosharp_b200.train import TrainStepV8  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
size = sys.argv[2] if len(sys.argv) > 2 else "n"
m = oracle_model("v8", "detect", size)
ts = TrainStepV8(m.state_dict(), size, 80, device="cuda")
x = synth_image(B, 640, 640).cuda()
t = _targets(B)
losses = []
for i in range(4):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    items = ts.step(x, t)
    b.record()
    torch.cuda.synchronize()
    losses.append([round(float(v), 4) for v in items])
    print(f"step {i}: {a.elapsed_time(b):.1f} ms  loss items {losses[-1]}", flush=True)
print(f"YOLOv8{size} B={B} 640x640 fp32 parity path: {B / (a.elapsed_time(b) / 1e3):.1f} images/s per training step")
