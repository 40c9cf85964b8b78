"""Test-only stand-in for yolosharp_b200.train.KernelOps: the same interface on PyTorch CPU ops, so that the GRAPH
logic of TrainStepV8 (wiring, concat / chunk / residual / pool / upsample backward, parameter bookkeeping) can be
checked against autograd through the oracle model without a GPU.  The kernels themselves are checked one by one in
tests/test_gpu_parity.py; the -m gpu test of the whole step swaps the real KernelOps in."""
import torch
import torch.nn.functional as F

from oracle import loss as oloss


class TorchOps:
    def conv_forward(self, x, w, bias, stride, pad):
        return F.conv2d(x.permute(0, 3, 1, 2), w, bias, stride, pad).permute(0, 2, 3, 1).contiguous()

    def conv_backward(self, x, dz, w, stride, pad, need_dx=True):
        xn, dzn = x.permute(0, 3, 1, 2), dz.permute(0, 3, 1, 2)
        dx = torch.nn.grad.conv2d_input(xn.shape, w, dzn, stride, pad)
        dw = torch.nn.grad.conv2d_weight(xn, w.shape, dzn, stride, pad)
        return dx.permute(0, 2, 3, 1).contiguous(), dw

    def bn_silu_forward(self, z, gamma, beta, rm, rv, act):
        zn = z.permute(0, 3, 1, 2)
        mean = zn.mean((0, 2, 3))
        var = zn.var((0, 2, 3), unbiased=False)
        u = F.batch_norm(zn, rm, rv, gamma, beta, training=True, momentum=0.03, eps=1e-3)
        y = F.silu(u) if act else u
        return y.permute(0, 2, 3, 1).contiguous(), mean, 1.0 / torch.sqrt(var + 1e-3)

    def bn_silu_backward(self, z, dy, gamma, beta, mean, invstd, act):
        zn = z.permute(0, 3, 1, 2).detach().clone().requires_grad_(True)
        g, b = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
        u = F.batch_norm(zn, None, None, g, b, training=True, momentum=0.03, eps=1e-3)
        y = F.silu(u) if act else u
        dz, dg, db = torch.autograd.grad(y, (zn, g, b), dy.permute(0, 3, 1, 2))
        return dz.permute(0, 2, 3, 1).contiguous(), dg, db

    def detection_loss(self, boxes, scores, targets, H, W):
        crit = oloss.V8DetectionLoss(scores.shape[1])
        b, s = boxes.detach().clone().requires_grad_(True), scores.detach().clone().requires_grad_(True)
        hw = [(H // st, W // st) for st in (8, 16, 32)]
        feats = [torch.zeros(1, 1, h, w) for h, w in hw]
        t = torch.as_tensor(targets, dtype=torch.float32).reshape(-1, 6)
        batch = {"batch_idx": t[:, 0], "cls": t[:, 1], "bboxes": t[:, 2:]}
        loss, items = crit({"boxes": b, "scores": s, "feats": feats}, batch)
        gb, gs = torch.autograd.grad(loss.sum(), (b, s), allow_unused=True)  # no targets: the box branch is unused
        return items, gb if gb is not None else torch.zeros_like(b), gs if gs is not None else torch.zeros_like(s)

    def adamw(self, p, g, m, v, step, lr, wd):
        b1, b2, eps = 0.9, 0.999, 1e-8
        p.mul_(1 - lr * wd)
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / (1 - b2 ** step) ** 0.5).add_(eps)
        p.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))

    # ---- v11 extras (no library kernels yet: tests/test_train_step.py checks the graph logic only) ----
    def gconv_forward(self, x, w, stride, pad, groups):
        return F.conv2d(x.permute(0, 3, 1, 2), w, None, stride, pad, 1, groups).permute(0, 2, 3, 1).contiguous()

    def gconv_backward(self, x, dz, w, stride, pad, groups):
        xn, dzn = x.permute(0, 3, 1, 2), dz.permute(0, 3, 1, 2)
        dx = torch.nn.grad.conv2d_input(xn.shape, w, dzn, stride, pad, 1, groups)
        dw = torch.nn.grad.conv2d_weight(xn, w.shape, dzn, stride, pad, 1, groups)
        return dx.permute(0, 2, 3, 1).contiguous(), dw

    @staticmethod
    def _attn(q, k, v, scale):
        # q, k (B, N, nh, kd), v (B, N, nh, hd): out[b, n, h] = sum_m softmax_m(q_n . k_m * scale) v_m
        a = torch.einsum("bnhd,bmhd->bhnm", q, k) * scale
        return torch.einsum("bhnm,bmhd->bnhd", a.softmax(-1), v)

    def attention_forward(self, q, k, v, scale):
        return self._attn(q, k, v, scale)

    def attention_backward(self, q, k, v, scale, dout):
        q, k, v = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        return torch.autograd.grad(self._attn(q, k, v, scale), (q, k, v), dout)
