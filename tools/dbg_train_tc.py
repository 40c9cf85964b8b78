"""Per-layer difference between the TF32 tensor-core training forward and the fp32 parity kernels (same weights, same
batch): conv outputs z of every Conv block in network order.  python tools/dbg_train_tc.py [v8|v11] [H] [W] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests.util import oracle_model, synth_image  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "v8"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 64
W = int(sys.argv[3]) if len(sys.argv) > 3 else 96
B = int(sys.argv[4]) if len(sys.argv) > 4 else 2
if arch == "v8":
    from yolosharp_b200.train import KernelOps as Ops, TrainStepV8 as Step
else:
    from yolosharp_b200.train_v11 import KernelOpsV11 as Ops, TrainStepV11 as Step
torch.manual_seed(0)
m = oracle_model(arch, "detect", "n")
sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
x = synth_image(B, H, W).cuda()
a = Step(sd0, "n", 80, device="cuda", ops=Ops(tensor_cores=False))
b = Step(sd0, "n", 80, device="cuda", ops=Ops(tensor_cores=True))
oa, ob = a.forward(x), b.forward(x)


def walk(o, seen, out):
    if id(o) in seen or isinstance(o, (torch.Tensor, str, int, float, type(None))):
        return
    seen.add(id(o))
    if hasattr(o, "z") and hasattr(o, "name") and isinstance(getattr(o, "z"), torch.Tensor):
        out.append(o)
    if isinstance(o, (list, tuple)):
        for e in o:
            walk(e, seen, out)
    elif hasattr(o, "__dict__") and not isinstance(o, (Step, Ops)):
        for k, v in vars(o).items():
            if k not in ("net", "ops", "P"):
                walk(v, seen, out)


la, lb = [], []
walk(a.layers, set(), la); walk(a.detect, set(), la)
walk(b.layers, set(), lb); walk(b.detect, set(), lb)
print(f"# {arch} {B}x{H}x{W}: {len(la)} conv blocks; columns: name, z shape, rms rel err, max abs / rms")
for p, q in zip(la, lb):
    assert p.name == q.name
    d = (q.z - p.z)
    rms = p.z.pow(2).mean().sqrt().clamp_min(1e-20)
    print(f"{p.name:28s} {str(tuple(p.z.shape)):22s} {float(d.pow(2).mean().sqrt() / rms):.2e} {float(d.abs().max() / rms):.2e}")
for u, v, n in ((oa[0], ob[0], "boxes"), (oa[1], ob[1], "scores")):
    print(n, float((v - u).pow(2).mean().sqrt() / u.pow(2).mean().sqrt()), float((v - u).abs().max() / u.pow(2).mean().sqrt()))

# ---- backward: both paths driven by the fp32 path's loss gradient ----
from bench import synth_targets  # noqa: E402
targets = synth_targets(B, 5)
items, gb, gs = a.ops.detection_loss(oa[0], oa[1], targets, H, W)
a.P.grad.zero_(); b.P.grad.zero_()
a.backward(gb, gs); b.backward(gb, gs)
ga, gt = a.P.grad.double(), b.P.grad.double()
print(f"# flat gradient: rel L2 {float((gt - ga).norm() / ga.norm()):.3e} cosine {float((gt * ga).sum() / (gt.norm() * ga.norm())):.6f}")
print("# per conv weight (network order): rel L2 of the gradient tensor, its share of the flat gradient norm")
tot = float(ga.norm())
for p in la:
    for suffix in (".conv.weight",):
        k = p.name + suffix
        if k in a.P.off:
            u, v = a.P.g(k).double(), b.P.g(k).double()
            print(f"{k:40s} {float((v - u).norm() / u.norm().clamp_min(1e-30)):.2e} {float(u.norm()) / tot:.2e}")
