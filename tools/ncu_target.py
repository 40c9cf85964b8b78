"""Target for ncu captures: two eager forwards (no CUDA graph) so that kernel launch indices map to
op indices.  python tools/ncu_target.py v8n 32"""
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
e = y.Engine(arch, size, task, 80, "f16", 0, B, 640, 640, flags=2)
e.load_state_dict(m.state_dict())
e.finalize()
in_dt = {"f16": torch.float16, "u8": torch.uint8, "f32": torch.float32}[os.environ.get("YB_IN_DTYPE", "f16")]
x = synth_image(B, 640, 640, dtype=in_dt).cuda()
names = e.op_names()
tc = [n for n in names if "decode" not in n and n != "model.0" and not n.endswith(".m") and n not in ("model.10", "model.13")]
if len(sys.argv) > 3:
    print("conv_tc launch order:", {n: i for i, n in enumerate(tc)})
for _ in range(2):
    e.forward(x)
torch.cuda.synchronize()
y.nms(e.forward(x), 0.25, 0.45)
torch.cuda.synchronize()
