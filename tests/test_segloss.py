"""The instance-mask term of v8SegmentationLoss (Utils/Loss.cs:688-865): oracle properties on the CPU, the CUDA entry point
yb_segmentation_loss against autograd through the oracle on the GPU."""
import pytest
import torch

from oracle import loss as oloss


def _case(B=2, A=300, nm=32, mh=40, mw=48, n_inst=5, p_fg=0.1, seed=0, img=(160.0, 192.0)):
    g = torch.Generator().manual_seed(seed)
    fg = torch.rand(B, A, generator=g) < p_fg
    gt_idx = torch.randint(0, n_inst, (B, A), generator=g)
    H, W = img
    # instance rectangles painted into the overlap-encoded mask (later instances on top), boxes = those rectangles
    masks = torch.zeros(B, mh, mw)
    inst_box = torch.zeros(B, n_inst, 4)
    for b in range(B):
        for i in range(n_inst):
            x1, y1 = float(torch.rand(1, generator=g)) * 0.6 * W, float(torch.rand(1, generator=g)) * 0.6 * H
            w, h = (0.08 + 0.3 * float(torch.rand(1, generator=g))) * W, (0.08 + 0.3 * float(torch.rand(1, generator=g))) * H
            inst_box[b, i] = torch.tensor([x1, y1, x1 + w, y1 + h])
            c0, c1, r0, r1 = int(x1 / W * mw), int((x1 + w) / W * mw) + 1, int(y1 / H * mh), int((y1 + h) / H * mh) + 1
            masks[b, r0:r1, c0:c1] = i + 1
    tb = torch.gather(inst_box, 1, gt_idx.unsqueeze(-1).expand(B, A, 4)) + torch.randn(B, A, 4, generator=g) * 0.5
    proto = torch.randn(B, nm, mh, mw, generator=g)
    coef = torch.randn(B, nm, A, generator=g) * 0.5
    return fg, gt_idx, tb, masks, proto, coef, torch.tensor([H, W])


def test_oracle_mask_loss_properties():
    fg, gi, tb, masks, proto, coef, imgsz = _case(A=60, mh=20, mw=24)
    scaled, item = oloss.segmentation_mask_loss(fg, gi, tb, masks, proto, coef, imgsz)
    assert torch.isfinite(item) and abs(float(scaled) - 2 * float(item)) < 1e-5  # loss * batch_size
    # logits that reproduce the ground truth drive the loss to ~0: proto channel 0 = +-20 per instance is not expressible
    # for several instances at once, so check the single-instance case
    fg1, gi1 = torch.zeros_like(fg), torch.zeros_like(gi)
    fg1[0, 0] = True
    gt = (masks[0] == 1).float()
    proto1 = torch.zeros_like(proto)
    proto1[0, 0] = (gt * 2 - 1) * 20
    coef1 = torch.zeros_like(coef)
    coef1[0, 0, 0] = 1.0
    _, item1 = oloss.segmentation_mask_loss(fg1, gi1, tb, masks, proto1, coef1, imgsz)
    assert float(item1) < 1e-6
    # no foreground anchors: zero loss, zero gradients
    p0, c0 = proto.clone().requires_grad_(True), coef.clone().requires_grad_(True)
    l0, _ = oloss.segmentation_mask_loss(torch.zeros_like(fg), gi, tb, masks, p0, c0, imgsz)
    l0.backward()
    assert float(l0) == 0.0 and float(p0.grad.abs().max()) == 0.0 and float(c0.grad.abs().max()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(B=3, A=2100, mh=80, mw=80, n_inst=9, p_fg=0.06, seed=1, img=(320.0, 320.0)),
                                dict(B=1, A=50, nm=8, mh=16, mw=16, p_fg=0.3, seed=2, img=(64.0, 64.0)), dict(p_fg=0.0, seed=3)])
def test_segmentation_loss_matches_oracle_gpu(kw):
    from yolosharp_b200 import engine as E
    fg, gi, tb, masks, proto, coef, imgsz = _case(**kw)
    p, c = proto.clone().requires_grad_(True), coef.clone().requires_grad_(True)
    scaled, item = oloss.segmentation_mask_loss(fg, gi, tb, masks, p, c, imgsz)
    scaled.backward()
    out = E.segmentation_loss(fg.cuda(), gi.cuda(), tb.cuda(), masks.cuda(), proto.cuda().contiguous(), coef.cuda().contiguous(), float(imgsz[0]),
                              float(imgsz[1]))
    torch.cuda.synchronize()
    torch.testing.assert_close(out["item"].cpu().view(()), item.float().view(()), rtol=2e-5, atol=1e-6)
    for got, ref, name in ((out["grad_coefficient"], c.grad, "coefficient"), (out["grad_proto"], p.grad, "proto")):
        scale = float(ref.abs().max()) or 1.0
        err = float((got.cpu() - ref).abs().max()) / scale
        assert err < 2e-5, f"grad {name}: max scaled error {err:.3e}"


def test_oracle_reproduces_golden_mask_loss():
    """tests/golden/r2_val_seg.npz (make_golden_r2b.py): the committed segmentation-mask-loss vectors."""
    import os

    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "r2_val_seg.npz"))
    t = lambda k: torch.from_numpy(g[k])
    p, c = t("seg_proto").clone().requires_grad_(True), t("seg_coef").clone().requires_grad_(True)
    scaled, item = oloss.segmentation_mask_loss(t("seg_fg"), t("seg_gt_idx"), t("seg_tbox"), t("seg_masks"), p, c, t("seg_imgsz"))
    scaled.backward()
    assert abs(float(item) - float(g["seg_item"])) < 1e-6 * max(1.0, abs(float(g["seg_item"])))
    torch.testing.assert_close(p.grad, t("seg_grad_proto"), rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(c.grad, t("seg_grad_coef"), rtol=1e-5, atol=1e-8)
