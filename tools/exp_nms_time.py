"""NMS kernel time on the bench's own prediction tensor + candidate statistics.  python tools/exp_nms_time.py [model] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import yolosharp_b200 as y  # noqa: E402
from bench import MODELS  # noqa: E402
from tests.util import oracle_model, synth_image  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "v8n"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
arch, size, task, _ = MODELS[model]
m = oracle_model(arch, task, size)
e = y.Engine(arch, size, task, 80, "f16", 0, B, 640, 640)
e.load_state_dict(m.state_dict())
e.finalize()
x = synth_image(B, 640, 640, dtype=torch.float16).cuda()
pred = e.forward(x)
if isinstance(pred, tuple):
    pred = pred[0]
torch.cuda.synchronize()
conf, cls = pred[:, 4:84].max(1)
cand = conf > 0.25
n = cand.sum(1)
print("candidates per image: min/mean/max", int(n.min()), float(n.float().mean()), int(n.max()))
for i in range(min(B, 3)):
    c = torch.bincount(cls[i][cand[i]], minlength=80)
    print(" image", i, "n", int(n[i]), "top class counts", sorted(c.tolist(), reverse=True)[:6])
out = y.nms(pred, 0.25, 0.45)
print("kept per image: mean", float(out[1].float().mean()))
for _ in range(3):
    y.nms(pred, 0.25, 0.45, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    y.nms(pred, 0.25, 0.45, out=out)
b.record()
torch.cuda.synchronize()
print(f"yb_nms B={B}: {a.elapsed_time(b) / 20:.4f} ms per call (general path forced: {bool(os.environ.get('YB_DEBUG_NMS_GENERAL'))})")
