import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolosharp_b200 as y
from tests.util import oracle_model, synth_image, oracle_activations, rel_err
from tests.test_gpu_parity import match_detections
import oracle.ops as oops
m = oracle_model("v8", "detect", "n")
x = synth_image(4, 640, 640)
u8 = synth_image(4, 640, 640, dtype=torch.uint8)
e = y.Engine("v8", "n", "detect", 80, "f16", 0, 4, 640, 640, flags=2 | 8)
e.load_state_dict(m.state_dict()); e.finalize()
(_, _), acts = oracle_activations(m, x)
ref0 = acts["model.0"]
for name, inp in (("f16", x.half()), ("f32", x), ("u8", u8)):
    p = e.forward(inp.cuda()); torch.cuda.synchronize()
    got = e.read_activation(0, 4)
    d = (got.float() - ref0).abs()
    print(name, "stem max err", float(d.max()), "border cols", float(d[..., 0].max()), float(d[..., -1].max()),
          "border rows", float(d[:, :, 0].max()), float(d[:, :, -1].max()), "rel", rel_err(got, ref0))
with torch.no_grad():
    ref = m(x)[0]["boxes"]
pred = e.forward(x.half().cuda())
out, keep = y.nms(pred, 0.25, 0.45)[:2] if False else (None, None)
from yolosharp_b200 import api
o, k = api.Ops.non_max_suppression(pred, 0.25, 0.45)
oref, _ = oops.non_max_suppression(ref, 0.25, 0.45)
for i in range(4):
    strong = oref[i][oref[i][:, 4] > 0.35]
    print("img", i, "strong", len(strong), "match", match_detections(strong, o[i].cpu(), iou_thr=0.85))
# timing u8 vs f16 input, B=32
e32 = y.Engine("v8", "n", "detect", 80, "f16", 0, 32, 640, 640)
e32.load_state_dict(m.state_dict()); e32.finalize()
for name, inp in (("f16", synth_image(32, 640, 640, dtype=torch.float16)), ("u8", synth_image(32, 640, 640, dtype=torch.uint8)), ("f32", synth_image(32, 640, 640))):
    xi = inp.cuda()
    outp = torch.empty((32, e32.pred_channels, e32.anchors), dtype=torch.float32, device="cuda")
    for _ in range(5): e32.forward(xi, out_pred=outp)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(30): e32.forward(xi, out_pred=outp)
    b.record(); torch.cuda.synchronize()
    print(name, "forward ms", a.elapsed_time(b) / 30)
