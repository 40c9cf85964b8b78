import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolosharp_b200 as y
from tests.util import oracle_model, synth_image
size, B, H, W = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
m = oracle_model("v8", "detect", size)
x = synth_image(B, H, W)
e = y.Engine("v8", size, "detect", 80, "f16", 0, B, H, W, flags=int(os.environ.get("YBF", "0")))
e.load_state_dict(m.state_dict()); e.finalize()
xc = x.cuda()
for _ in range(3):
    p = e.forward(xc); torch.cuda.synchronize()
with torch.no_grad(): ref = m(x)[0]["boxes"]
print(size, "max err", float((p.cpu() - ref).abs().max()))
