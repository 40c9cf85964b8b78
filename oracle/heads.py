"""Oracle restatements (test infrastructure only) of the decode tails of the OBB and Pose heads and of the rotated NMS.
Every function cites the reference lines it follows (paths relative to /root/reference/YoloSharp)."""
import math

import torch


def dfl_expectation(box_logits, reg_max=16):
    """Block.cs:15-45 `DFL.forward`: (B, 4 * reg_max, A) -> (B, 4, A): softmax over the bins, conv with weights arange."""
    b, _, a = box_logits.shape
    x = box_logits.view(b, 4, reg_max, a).transpose(2, 1).softmax(1)          # Block.cs:41-43
    w = torch.arange(reg_max, dtype=torch.float32).view(1, reg_max, 1, 1)     # Block.cs:29-30
    return (x * w).sum(1)                                                     # the 1x1 conv with those weights, :44


def dist2rbox(pred_dist, pred_angle, anchor_points, dim=-1):
    """Utils/Tal.cs:389-408."""
    lt, rb = pred_dist.split(2, dim=dim)
    cos, sin = torch.cos(pred_angle), torch.sin(pred_angle)
    xf, yf = ((rb - lt) / 2).split(1, dim=dim)
    x, y = xf * cos - yf * sin, xf * sin + yf * cos
    xy = torch.cat([x, y], dim=dim) + anchor_points
    return torch.cat([xy, lt + rb], dim=dim)


def obb_inference(box_logits, cls_logits, angle_logits, anchors, strides, reg_max=16):
    """Modules/Head.cs:410-436 (`Obb._inference` / `forward_head` / `decode_bboxes`) on top of Head.cs:204-223:
    angle = (sigmoid - 0.25) * pi, dbox = dist2rbox(dfl(boxes), angle, anchors, dim: 1) * strides,
    y = cat(dbox, cls.sigmoid(), angle) of shape (B, 4 + nc + 1, A).  anchors (2, A), strides (A)."""
    angle = (angle_logits.sigmoid() - 0.25) * math.pi                         # Head.cs:428
    dbox = dist2rbox(dfl_expectation(box_logits, reg_max), angle, anchors.unsqueeze(0), dim=1) * strides   # :436, :221
    return torch.cat([dbox, cls_logits.sigmoid(), angle], 1)                  # Head.cs:222, :416


def kpts_decode(kpts, anchors, strides, ndim=3):
    """Modules/Head.cs:595-609: kpts (B, nk, A)."""
    y = kpts.clone()
    if ndim == 3:
        y[:, 2::ndim] = y[:, 2::ndim].sigmoid()
    y[:, 0::ndim] = (y[:, 0::ndim] * 2.0 + (anchors[0] - 0.5)) * strides
    y[:, 1::ndim] = (y[:, 1::ndim] * 2.0 + (anchors[1] - 0.5)) * strides
    return y


def _get_covariance_matrix(boxes):
    """Utils/Metrics.cs:260-280."""
    gbbs = torch.cat([boxes[..., 2:4].pow(2) / 12, boxes[..., 4:]], dim=-1)
    a, b, c = gbbs.split(1, dim=-1)
    cos, sin = c.cos(), c.sin()
    cos2, sin2 = cos.pow(2), sin.pow(2)
    return a * cos2 + b * sin2, a * sin2 + b * cos2, (a - b) * cos * sin


def batch_probiou(obb1, obb2, eps=1e-7):
    """Utils/Metrics.cs:223-254: (N, 5), (M, 5) xywhr -> (N, M)."""
    x1, y1 = obb1[..., 0].unsqueeze(-1), obb1[..., 1].unsqueeze(-1)
    x2, y2 = obb2[..., 0].unsqueeze(0), obb2[..., 1].unsqueeze(0)
    a1, b1, c1 = _get_covariance_matrix(obb1)
    a2, b2, c2 = (t.squeeze(-1)[None] for t in _get_covariance_matrix(obb2))
    t1 = (((a1 + a2) * (y1 - y2).pow(2) + (b1 + b2) * (x1 - x2).pow(2)) / ((a1 + a2) * (b1 + b2) - (c1 + c2).pow(2) + eps)) * 0.25
    t2 = (((c1 + c2) * (x2 - x1) * (y1 - y2)) / ((a1 + a2) * (b1 + b2) - (c1 + c2).pow(2) + eps)) * 0.5
    t3 = (((a1 + a2) * (b1 + b2) - (c1 + c2).pow(2))
          / (4 * ((a1 * b1 - c1.pow(2)).clamp_(0) * (a2 * b2 - c2.pow(2)).clamp_(0)).sqrt() + eps) + eps).log() * 0.5
    bd = (t1 + t2 + t3).clamp(eps, 100.0)
    hd = (1.0 - (-bd).exp() + eps).sqrt()
    return 1 - hd


def nms_rotated(boxes, scores, threshold=0.45):
    """Utils/Ops.cs:373-401, use_triu branch."""
    sorted_idx = torch.argsort(scores, descending=True)
    boxes = boxes[sorted_idx]
    ious = batch_probiou(boxes, boxes).triu_(diagonal=1)
    pick = torch.nonzero((ious >= threshold).sum(0) <= 0).squeeze_(-1)
    return sorted_idx[pick]
