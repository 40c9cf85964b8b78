import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolosharp_b200 as y
from tests.util import oracle_model, synth_image
m = oracle_model("v8", "detect", "n")
x = synth_image(1, 64, 64)
e = y.Engine("v8", "n", "detect", 80, "f16", 0, 1, 64, 64, flags=2 | 8)
e.load_state_dict(m.state_dict()); e.finalize()
p = e.forward(x.cuda()); torch.cuda.synchronize()
with torch.no_grad(): ref = m(x)[0]["boxes"]
print("max err", float((p.cpu() - ref).abs().max()))
got = e.read_activation(0, 1)
from tests.util import oracle_activations, rel_err
(_, _), acts = oracle_activations(m, x)
print("stem rel err", rel_err(got, acts["model.0"]))
