import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from tests.test_train_step import _targets
from tests.util import oracle_model, synth_image
from yolosharp_b200.train_native import NativeTrainer
torch.manual_seed(0)
m = oracle_model("v8", "detect", "n")
sd = m.state_dict()
u8 = synth_image(2, 64, 96, dtype=torch.uint8).cuda()
xf = u8.float() / 255.0
def run(x, keep=[]):
    t = NativeTrainer(sd, "v8", "n", 80, device="cuda", max_batch=2, height=64, width=96, lr=1e-3)
    it = t.step(x, _targets(2))
    keep.append(t)
    return it, t.grad.clone()
a, ga = run(u8); b, gb = run(xf); c, gc = run(xf); d, gd = run(u8)
print("u8  ", a.tolist()); print("f32 ", b.tolist()); print("f32 ", c.tolist()); print("u8  ", d.tolist())
print("grad u8-f32", float((ga-gb).norm()/gb.norm()), "f32-f32", float((gb-gc).norm()/gb.norm()), "u8-u8", float((ga-gd).norm()/ga.norm()))
print("synth f32 vs u8/255:", float((synth_image(2,64,96).cuda() - xf).abs().max()))
