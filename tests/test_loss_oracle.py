"""CPU tests of oracle/loss.py (the checker of the training-path kernels).
The reference ships no tests or golden values for its loss, so the oracle is pinned by (1) an independent
scalar (python-loop) restatement of the assigner on small cases, (2) closed-form properties of the formulas,
(3) autograd consistency (finite differences) - the gradients are what the CUDA kernels will be checked against."""
import math

import numpy as np
import pytest
import torch

from oracle import loss as oloss
from tests.util import oracle_model, synth_image


def _scalar_ciou(b1, b2, eps=1e-7):
    """Metrics.cs:36-111 (xywh=false, CIoU) with python floats."""
    x1, y1, x2, y2 = [float(v) for v in b1]
    X1, Y1, X2, Y2 = [float(v) for v in b2]
    w1, h1 = x2 - x1, max(y2 - y1, eps)
    w2, h2 = X2 - X1, max(Y2 - Y1, eps)
    inter = max(min(x2, X2) - max(x1, X1), 0.0) * max(min(y2, Y2) - max(y1, Y1), 0.0)
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    cw, ch = max(x2, X2) - min(x1, X1), max(y2, Y2) - min(y1, Y1)
    c2 = cw * cw + ch * ch + eps
    rho2 = ((X1 + X2 - x1 - x2) ** 2 + (Y1 + Y2 - y1 - y2) ** 2) / 4
    v = 4 / math.pi ** 2 * (math.atan(w2 / h2) - math.atan(w1 / h1)) ** 2
    alpha = v / (v - iou + (1 + eps))
    return iou - (rho2 / c2 + v * alpha)


def _scalar_assigner(pd_scores, pd_bboxes, anc, gt_labels, gt_bboxes, mask_gt, topk, nc, alpha, beta, stride):
    """Loop restatement of Tal.cs:70-266 for ONE image; returns fg mask, gt index, target scores."""
    A, n = anc.shape[0], gt_bboxes.shape[0]
    stride_val = stride[1]
    in_gts = np.zeros((n, A), bool)
    overlaps = np.zeros((n, A))
    metric = np.zeros((n, A))
    for g in range(n):
        x1, y1, x2, y2 = [float(v) for v in gt_bboxes[g]]
        cx, cy, w, h = (x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1
        if mask_gt[g] and w < stride[0]:
            w = stride_val
        if mask_gt[g] and h < stride[0]:
            h = stride_val
        bx1, by1, bx2, by2 = cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2
        for a in range(A):
            ax, ay = float(anc[a, 0]), float(anc[a, 1])
            in_gts[g, a] = min(ax - bx1, ay - by1, bx2 - ax, by2 - ay) > 1e-9
            if in_gts[g, a] and mask_gt[g]:
                ov = max(_scalar_ciou(gt_bboxes[g], pd_bboxes[a]), 0.0)
                overlaps[g, a] = ov
                metric[g, a] = float(pd_scores[a, int(gt_labels[g])]) ** alpha * ov ** beta
    mask_pos = np.zeros((n, A), bool)
    for g in range(n):
        if not mask_gt[g]:
            continue
        top = np.argsort(-metric[g], kind="stable")[:topk]
        for a in top:
            mask_pos[g, a] = in_gts[g, a]
    fg = mask_pos.sum(0)
    for a in range(A):
        if fg[a] > 1:
            best = int(np.argmax(overlaps[:, a]))
            mask_pos[:, a] = False
            mask_pos[best, a] = True
    fg = mask_pos.sum(0) > 0
    gt_idx = mask_pos.argmax(0)
    target_scores = np.zeros((A, nc))
    am = metric * mask_pos
    pos_am = am.max(1, keepdims=True)
    pos_ov = (overlaps * mask_pos).max(1, keepdims=True)
    norm = (am * pos_ov / (pos_am + 1e-9)).max(0)
    for a in range(A):
        if fg[a]:
            target_scores[a, int(gt_labels[gt_idx[a]])] = norm[a]
    return fg, gt_idx, target_scores


def _random_case(seed, A_hw=((8, 8), (4, 4), (2, 2)), n=5, nc=4):
    g = torch.Generator().manual_seed(seed)
    strides = [8, 16, 32]
    feats = [torch.zeros(1, 1, h, w) for h, w in A_hw]
    anc, st = oloss.make_anchors(feats, strides, 0.5)
    A = anc.shape[0]
    img = A_hw[0][0] * 8
    centers = torch.rand(n, 2, generator=g) * img
    wh = torch.rand(n, 2, generator=g) * img * 0.5 + 3
    gt = torch.cat((centers - wh / 2, centers + wh / 2), 1).clamp(0, img)
    labels = torch.randint(0, nc, (n, 1), generator=g).float()
    pd_scores = torch.rand(1, A, nc, generator=g)
    jitter = torch.randn(A, 4, generator=g) * 6
    pix = anc * st
    pd_boxes = torch.cat((pix - 12, pix + 12), 1) + jitter
    return anc, st, gt, labels, pd_scores, pd_boxes[None], nc, strides


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_assigner_matches_scalar_restatement(seed):
    anc, st, gt, labels, pd_scores, pd_boxes, nc, strides = _random_case(seed)
    mask_gt = torch.ones(1, gt.shape[0], 1)
    mask_gt[0, -1] = 0  # one padded (invalid) ground truth
    ta = oloss.TaskAlignedAssigner(topk=10, num_classes=nc, alpha=0.5, beta=6.0, stride=strides)
    _, tb, ts, fg, gi = ta.forward(pd_scores, pd_boxes, anc * st, labels[None], gt[None], mask_gt)
    fg2, gi2, ts2 = _scalar_assigner(pd_scores[0].numpy(), pd_boxes[0].numpy(), (anc * st).numpy(), labels[:, 0].numpy(),
                                     gt.numpy(), mask_gt[0, :, 0].numpy() > 0, 10, nc, 0.5, 6.0, strides)
    assert fg.sum() > 0
    np.testing.assert_array_equal(fg[0].numpy(), fg2)
    np.testing.assert_array_equal(gi[0].numpy()[fg2], gi2[fg2])
    np.testing.assert_allclose(ts[0].numpy(), ts2, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(tb[0].numpy()[fg2], gt.numpy()[gi2[fg2]])


def test_ciou_properties_and_scalar():
    g = torch.Generator().manual_seed(0)
    c = torch.rand(64, 2, generator=g) * 100
    wh = torch.rand(64, 2, generator=g) * 40 + 1
    b1 = torch.cat((c - wh / 2, c + wh / 2), 1)
    b2 = b1 + torch.randn(64, 4, generator=g) * 5
    out = oloss.bbox_iou_ciou(b1, b2)
    ref = np.array([_scalar_ciou(b1[i], b2[i]) for i in range(64)])
    np.testing.assert_allclose(out[:, 0].numpy(), ref, rtol=1e-4, atol=1e-5)
    same = oloss.bbox_iou_ciou(b1, b1)
    assert float((same - 1).abs().max()) < 1e-5       # identical boxes -> CIoU = IoU = 1
    assert float(out.max()) <= 1.0 + 1e-6


def test_dfl_is_soft_cross_entropy():
    """DFLoss (Loss.cs:104-119): for an integer target t the loss is CE(pred, t); the gradient of the
    expectation target t+0.5 is the average of the two neighbouring CE gradients."""
    pred = torch.randn(8, 16)
    t = torch.tensor([[3.0, 7.0], [0.0, 14.0], [5.5, 2.25], [14.98, 1.0]])
    out = oloss.dfl_loss(pred, t)
    tl = t.clamp(0, 14.99).long()
    ce = torch.nn.functional.cross_entropy
    for i in range(4):
        for j in range(2):
            tt = float(t[i, j].clamp(0, 14.99))
            l, wl = int(tl[i, j]), int(tl[i, j]) + 1 - tt
            exp = ce(pred[i * 2 + j:i * 2 + j + 1], torch.tensor([l])) * wl + \
                  ce(pred[i * 2 + j:i * 2 + j + 1], torch.tensor([l + 1])) * (1 - wl)
            got_pair = out[i, 0] * 2 - sum(
                float(ce(pred[i * 2 + k:i * 2 + k + 1], torch.tensor([int(tl[i, k])])) * (int(tl[i, k]) + 1 - float(t[i, k].clamp(0, 14.99))) +
                      ce(pred[i * 2 + k:i * 2 + k + 1], torch.tensor([int(tl[i, k]) + 1])) * (1 - (int(tl[i, k]) + 1 - float(t[i, k].clamp(0, 14.99)))))
                for k in range(2) if k != j)
            assert abs(float(got_pair) - float(exp)) < 1e-4


def _head_outputs(B=2, H=96, W=128, nc=80):
    m = oracle_model("v8", "detect", "n", nc=nc).train()
    x = synth_image(B, H, W)
    with torch.no_grad():
        _, preds = m(x)
    return preds


def test_detection_loss_on_network_outputs_and_gradcheck():
    preds = _head_outputs()
    batch = {"batch_idx": torch.tensor([0, 0, 1, 1, 1]), "cls": torch.tensor([3, 17, 0, 3, 55]),
             "bboxes": torch.tensor([[0.5, 0.5, 0.4, 0.5], [0.25, 0.3, 0.2, 0.25], [0.7, 0.6, 0.3, 0.3],
                                     [0.3, 0.7, 0.5, 0.4], [0.52, 0.48, 0.06, 0.05]])}
    crit = oloss.V8DetectionLoss(80)
    boxes = preds["boxes"].clone().double().requires_grad_(True)
    scores = preds["scores"].clone().double().requires_grad_(True)
    p = {"boxes": boxes, "scores": scores, "feats": [f.double() for f in preds["feats"]]}
    crit.proj = crit.proj.double()
    loss, items = crit(p, batch)
    assert loss.shape == (3,) and torch.isfinite(loss).all() and (items > 0).all()
    np.testing.assert_allclose(loss.detach().numpy(), items.numpy() * 2)  # Loss.cs:473: loss * batch_size
    (targets, _) = crit.assigned_targets_and_loss(p, batch)
    fg, gt_idx, tbox, tscore = targets
    assert 10 <= int(fg.sum()) <= 50          # <= topk anchors per ground truth
    assert float(tscore.max()) <= 1.0 + 1e-6 and float(tscore.sum()) > 0
    # finite-difference check of d(sum loss)/d(inputs) for the FIXED assignment (the assigner runs under no_grad:
    # the dependence of the targets on the predictions is deliberately not part of the gradient)
    total = loss.sum()
    gb, gs = torch.autograd.grad(total, (boxes, scores))
    b_i, a_i = [int(v[0]) for v in torch.nonzero(fg, as_tuple=True)]
    eps = 1e-5
    for (tensor, grad, ch) in ((boxes, gb, 5), (boxes, gb, 40), (scores, gs, 3), (scores, gs, int(batch["cls"][0]))):
        base = tensor.detach().clone()
        vals = []
        for sgn in (+1, -1):
            t2 = base.clone()
            t2[b_i, ch, a_i] += sgn * eps
            q = dict(p)
            q["boxes" if tensor is boxes else "scores"] = t2
            vals.append(float(crit.loss_from_targets(q, targets).sum() * 2))
        fd = (vals[0] - vals[1]) / (2 * eps)
        assert abs(fd - float(grad[b_i, ch, a_i])) < 1e-4 * max(1.0, abs(fd)), (ch, fd, float(grad[b_i, ch, a_i]))


def test_no_targets_gives_pure_background_loss():
    preds = _head_outputs(B=1, H=64, W=64)
    crit = oloss.V8DetectionLoss(80)
    batch = {"batch_idx": torch.zeros(0), "cls": torch.zeros(0), "bboxes": torch.zeros(0, 4)}
    loss, items = crit(preds, batch)
    assert float(items[0]) == 0.0 and float(items[2]) == 0.0
    exp = torch.nn.functional.binary_cross_entropy_with_logits(preds["scores"], torch.zeros_like(preds["scores"]), reduction="sum") * 0.5
    assert abs(float(items[1]) - float(exp)) < 1e-3 * float(exp)
