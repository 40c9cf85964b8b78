"""Is the 1e-3 native-vs-Python gradient difference a property of the Python walk (ATen's atomicAdd max-pool backward)?
Two Python steps against each other, two native steps against each other, Python against native.  argv: arch H W"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests.test_train_step import _targets  # noqa: E402
from tests.util import oracle_model, synth_image  # noqa: E402
from yolosharp_b200.train_native import NativeTrainer  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "v8"
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (128, 160)
if arch == "v8":
    from yolosharp_b200.train import KernelOps as Ops, TrainStepV8 as Step
else:
    from yolosharp_b200.train_v11 import KernelOpsV11 as Ops, TrainStepV11 as Step
torch.manual_seed(0)
m = oracle_model(arch, "detect", "n")
sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
x, t = synth_image(2, H, W).cuda(), _targets(2)


def py():
    a = Step(sd0, "n", 80, device="cuda", ops=Ops(tensor_cores=True), lr=1e-3)
    a.step(x, t)
    return {k: a.P.g(k).double().clone() for k in a.P.names}


def nat():
    b = NativeTrainer(sd0, arch, "n", 80, device="cuda", max_batch=2, height=H, width=W, lr=1e-3)
    b.step(x, t)
    return {k: b.g(k).double().clone() for k in b.params}


def cmp(u, v, tag):
    fu, fv = torch.cat([u[k].reshape(-1) for k in u]), torch.cat([v[k].reshape(-1) for k in u])
    worst = max(((float((u[k] - v[k]).abs().max()) / max(float(u[k].abs().max()), 1e-3 * float(fu.abs().max())), k) for k in u))
    print(f"{arch} {H}x{W} {tag}: flat rel L2 {float((fu - fv).norm() / fu.norm()):.3e}  worst {worst[1]} {worst[0]:.2e}")


p1, p2, n1, n2 = py(), py(), nat(), nat()
cmp(p1, p2, "python vs python")
cmp(n1, n2, "native vs native")
cmp(p1, n1, "python vs native")

print("# python vs native per tensor (state-dict order), relative to the tensor's own max")
for k in p1:
    e = float((p1[k] - n1[k]).abs().max()) / max(float(p1[k].abs().max()), 1e-30)
    if k.endswith("conv.weight") or k.endswith(".weight") and ".bn." not in k:
        print(f"{k:44s} {e:.2e}")
