"""Reference point for the TF32 training kernels: the SAME comparison (head outputs and flat parameter gradient, TF32 vs
fp32) run with PyTorch's own CUDA convolutions (cuDNN, torch.backends.cudnn.allow_tf32 on / off) on the oracle model -
i.e. what the reference's libtorch back-end does on a GPU.  python tools/exp_torch_tf32.py [v8|v11] [H] [W] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import synth_targets  # noqa: E402
from oracle import loss as oloss  # noqa: E402
from tests.util import oracle_model, synth_image  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "v8"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 320
W = int(sys.argv[3]) if len(sys.argv) > 3 else 320
B = int(sys.argv[4]) if len(sys.argv) > 4 else 4
torch.manual_seed(0)
m = oracle_model(arch, "detect", "n").cuda().train()
x = synth_image(B, H, W).cuda()
t = synth_targets(B, 5).cuda()
batch = {"batch_idx": t[:, 0], "cls": t[:, 1], "bboxes": t[:, 2:]}
crit = oloss.V8DetectionLoss(80)
sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}


def run(tf32, head_grads=None):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    m.load_state_dict(sd0)
    m.zero_grad()
    _, preds = m(x)
    boxes, scores = preds["boxes"], preds["scores"]
    if head_grads is None:  # the oracle loss is CPU code: differentiate it on detached host copies of the head outputs
        bc, sc = boxes.detach().cpu().requires_grad_(True), scores.detach().cpu().requires_grad_(True)
        loss, _ = crit({"boxes": bc, "scores": sc, "feats": [f.detach().cpu() for f in preds["feats"]]},
                       {k: v.cpu() for k, v in batch.items()})
        gb, gs = torch.autograd.grad(loss.sum(), (bc, sc))
        gb, gs = gb.cuda(), gs.cuda()
    else:
        gb, gs = head_grads
    torch.autograd.backward((boxes, scores), (gb, gs))
    g = torch.cat([p.grad.reshape(-1) for k, p in m.named_parameters() if p.grad is not None]).double()
    return boxes.detach(), scores.detach(), g, (gb.detach(), gs.detach())


b0, s0, g0, hg = run(False)
b1, s1, g1, _ = run(True, hg)


def rel(u, v):
    return float((u - v).pow(2).mean().sqrt() / v.pow(2).mean().sqrt())


print(f"# torch/cuDNN {arch} {B}x{H}x{W}: TF32 vs fp32 head outputs rms rel box {rel(b1, b0):.2e} cls {rel(s1, s0):.2e}; "
      f"flat gradient rel L2 {float((g1 - g0).norm() / g0.norm()):.3e} cosine {float((g1 * g0).sum() / (g1.norm() * g0.norm())):.6f}")
