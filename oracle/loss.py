"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's detection training loss.

Restates, op for op, `Utils/Loss.cs` (v8DetectionLoss :328-485, BboxLoss :122-167, DFLoss :94-120),
`Utils/Tal.cs` (TaskAlignedAssigner :13-311, bbox2dist :364-378) and `Utils/Metrics.cs` (bbox_iou :36-111) of
IntptrMax/YoloSharp in PyTorch, so that autograd through it is the checker for the CUDA loss / gradient kernels of
the training path.  Parity is unpinned by reference tests (the reference ships
none and cannot run here); the anchors are its own call sites and the properties checked in tests/test_loss_oracle.py.

Two deliberate fidelity points (they differ from the Python Ultralytics code the C# was translated from):
  * bbox_iou keeps `alpha = v / (v - iou + (1 + eps))` INSIDE the autograd graph (Metrics.cs:101 has no no_grad
    scope), so the CIoU gradient includes d(alpha);
  * only the box HEIGHTS are clamped to eps (Metrics.cs:77-78), not offset by it.
Only tests/ and the CPU legs of bench.py may import this module.
"""
import math

import torch
import torch.nn.functional as F

from .modules import dist2bbox, make_anchors
from .ops import xywh2xyxy


def xyxy2xywh(x):
    """Utils/Ops.cs xyxy2xywh."""
    y = torch.empty_like(x)
    y[..., 0] = (x[..., 0] + x[..., 2]) / 2
    y[..., 1] = (x[..., 1] + x[..., 3]) / 2
    y[..., 2] = x[..., 2] - x[..., 0]
    y[..., 3] = x[..., 3] - x[..., 1]
    return y


def bbox_iou_ciou(box1, box2, eps=1e-7):
    """Metrics.cs:36-111 with xywh=false, CIoU=true.  Boxes (..., 4) xyxy -> (..., 1)."""
    b1_x1, b1_y1, b1_x2, b1_y2 = box1.chunk(4, -1)
    b2_x1, b2_y1, b2_x2, b2_y2 = box2.chunk(4, -1)
    w1, h1 = b1_x2 - b1_x1, (b1_y2 - b1_y1).clamp(eps)
    w2, h2 = b2_x2 - b2_x1, (b2_y2 - b2_y1).clamp(eps)
    inter = (torch.minimum(b1_x2, b2_x2) - torch.maximum(b1_x1, b2_x1)).clamp(0) * \
            (torch.minimum(b1_y2, b2_y2) - torch.maximum(b1_y1, b2_y1)).clamp(0)
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    cw = torch.maximum(b1_x2, b2_x2) - torch.minimum(b1_x1, b2_x1)
    ch = torch.maximum(b1_y2, b2_y2) - torch.minimum(b1_y1, b2_y1)
    c2 = cw.pow(2) + ch.pow(2) + eps
    rho2 = ((b2_x1 + b2_x2 - b1_x1 - b1_x2).pow(2) + (b2_y1 + b2_y2 - b1_y1 - b1_y2).pow(2)) / 4
    v = 4 / (math.pi * math.pi) * (torch.atan(w2 / h2) - torch.atan(w1 / h1)).pow(2)
    alpha = v / (v - iou + (1 + eps))  # NOT detached in the reference
    return iou - (rho2 / c2 + v * alpha)


def bbox2dist(anchor_points, bbox, reg_max=None):
    """Tal.cs:364-378."""
    x1y1, x2y2 = bbox.chunk(2, -1)
    dist = torch.cat((anchor_points - x1y1, x2y2 - anchor_points), -1)
    if reg_max is not None:
        dist = dist.clamp(0, reg_max - 0.01)
    return dist


class TaskAlignedAssigner:
    """Tal.cs:13-311 (non-rotated).  Inputs as in the reference:
    pd_scores (b, A, nc) sigmoid scores, pd_bboxes (b, A, 4) xyxy in pixels, anc_points (A, 2) pixels,
    gt_labels (b, n, 1), gt_bboxes (b, n, 4) xyxy pixels, mask_gt (b, n, 1)."""

    def __init__(self, topk=13, num_classes=80, alpha=1.0, beta=6.0, stride=None, eps=1e-9, topk2=None):
        self.topk, self.topk2 = topk, topk2 if topk2 is not None else topk
        self.num_classes, self.alpha, self.beta, self.eps = num_classes, alpha, beta, eps
        self.stride = stride or [8, 16, 32]
        self.stride_val = self.stride[1] if len(self.stride) > 1 else self.stride[0]

    @torch.no_grad()
    def forward(self, pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt):
        self.bs, self.n_max_boxes = pd_scores.shape[0], gt_bboxes.shape[1]
        if self.n_max_boxes == 0:  # Tal.cs:57-66
            return (torch.full_like(pd_scores[..., 0], self.num_classes), torch.zeros_like(pd_bboxes),
                    torch.zeros_like(pd_scores), torch.zeros_like(pd_scores[..., 0]), torch.zeros_like(pd_scores[..., 0]))
        mask_pos, align_metric, overlaps = self.get_pos_mask(pd_scores, pd_bboxes, gt_labels, gt_bboxes, anc_points, mask_gt)
        target_gt_idx, fg_mask, mask_pos = self.select_highest_overlaps(mask_pos, overlaps, align_metric)
        target_labels, target_bboxes, target_scores = self.get_targets(gt_labels, gt_bboxes, target_gt_idx, fg_mask)
        # normalise (Tal.cs:83-88)
        align_metric = align_metric * mask_pos
        pos_align_metrics = align_metric.amax(dim=-1, keepdim=True)
        pos_overlaps = (overlaps * mask_pos).amax(dim=-1, keepdim=True)
        norm_align_metric = (align_metric * pos_overlaps / (pos_align_metrics + self.eps)).amax(-2).unsqueeze(-1)
        target_scores = target_scores * norm_align_metric
        return target_labels, target_bboxes, target_scores, fg_mask.bool(), target_gt_idx

    def get_pos_mask(self, pd_scores, pd_bboxes, gt_labels, gt_bboxes, anc_points, mask_gt):
        mask_in_gts = self.select_candidates_in_gts(anc_points, gt_bboxes, mask_gt)
        align_metric, overlaps = self.get_box_metrics(pd_scores, pd_bboxes, gt_labels, gt_bboxes, mask_in_gts * mask_gt)
        mask_topk = self.select_topk_candidates(align_metric, mask_gt.expand(-1, -1, self.topk).bool())
        return mask_topk * mask_in_gts * mask_gt, align_metric, overlaps

    def get_box_metrics(self, pd_scores, pd_bboxes, gt_labels, gt_bboxes, mask_gt):
        na = pd_bboxes.shape[-2]
        mask_gt = mask_gt.bool()
        overlaps = torch.zeros(self.bs, self.n_max_boxes, na, dtype=pd_bboxes.dtype)
        bbox_scores = torch.zeros(self.bs, self.n_max_boxes, na, dtype=pd_scores.dtype)
        ind0 = torch.arange(self.bs).view(-1, 1).expand(-1, self.n_max_boxes)
        ind1 = gt_labels.squeeze(-1).long()
        bbox_scores[mask_gt] = pd_scores[ind0, :, ind1][mask_gt]
        pd_boxes = pd_bboxes.unsqueeze(1).expand(-1, self.n_max_boxes, -1, -1)[mask_gt]
        gt_boxes = gt_bboxes.unsqueeze(2).expand(-1, -1, na, -1)[mask_gt]
        overlaps[mask_gt] = bbox_iou_ciou(gt_boxes, pd_boxes).squeeze(-1).clamp(0)  # iou_calculation, Tal.cs:139-142
        return bbox_scores.pow(self.alpha) * overlaps.pow(self.beta), overlaps

    def select_topk_candidates(self, metrics, topk_mask):
        """Tal.cs:144-167: an anchor selected more than once (masked indices all point at 0) is dropped."""
        _, topk_idxs = torch.topk(metrics, self.topk, dim=-1, largest=True)
        topk_idxs = topk_idxs.masked_fill(~topk_mask, 0)
        count = torch.zeros(metrics.shape, dtype=torch.int8)
        ones = torch.ones_like(topk_idxs[:, :, :1], dtype=torch.int8)
        for k in range(self.topk):
            count.scatter_add_(-1, topk_idxs[:, :, k:k + 1], ones)
        count.masked_fill_(count > 1, 0)
        return count.to(metrics.dtype)

    def get_targets(self, gt_labels, gt_bboxes, target_gt_idx, fg_mask):
        batch_ind = torch.arange(self.bs)[..., None]
        target_gt_idx = target_gt_idx + batch_ind * self.n_max_boxes
        target_labels = gt_labels.long().flatten()[target_gt_idx].clamp(0)
        target_bboxes = gt_bboxes.view(-1, gt_bboxes.shape[-1])[target_gt_idx]
        target_scores = torch.zeros(target_labels.shape[0], target_labels.shape[1], self.num_classes, dtype=torch.int64)
        target_scores.scatter_(2, target_labels.unsqueeze(-1), 1)
        fg_scores_mask = fg_mask[:, :, None].repeat(1, 1, self.num_classes)
        target_scores = torch.where(fg_scores_mask > 0, target_scores, 0)
        return target_labels, target_bboxes, target_scores

    def select_candidates_in_gts(self, xy_centers, gt_bboxes, mask_gt, eps=1e-9):
        """Tal.cs:213-235: boxes narrower than the smallest stride are widened to stride_val first."""
        gt_xywh = xyxy2xywh(gt_bboxes)
        wh_mask = gt_xywh[..., 2:] < self.stride[0]
        gt_xywh[..., 2:] = torch.where((wh_mask * mask_gt).bool(),
                                       torch.tensor(float(self.stride_val), dtype=gt_xywh.dtype), gt_xywh[..., 2:])
        gt_bboxes = xywh2xyxy(gt_xywh)
        n_anchors, (bs, n_boxes) = xy_centers.shape[0], gt_bboxes.shape[:2]
        lt, rb = gt_bboxes.view(-1, 1, 4).chunk(2, 2)
        deltas = torch.cat((xy_centers[None] - lt, rb - xy_centers[None]), dim=2).view(bs, n_boxes, n_anchors, -1)
        return deltas.amin(3).gt(eps).to(gt_bboxes.dtype)

    def select_highest_overlaps(self, mask_pos, overlaps, align_metric):
        """Tal.cs:237-266."""
        fg_mask = mask_pos.sum(-2)
        if fg_mask.amax() > 1:
            mask_multi_gts = (fg_mask.unsqueeze(1) > 1).expand(self.bs, self.n_max_boxes, -1)
            max_overlaps_idx = overlaps.argmax(1)
            is_max = torch.zeros_like(mask_pos)
            is_max.scatter_(1, max_overlaps_idx.unsqueeze(1), 1)
            mask_pos = torch.where(mask_multi_gts, is_max, mask_pos).float()
            fg_mask = mask_pos.sum(-2)
        if self.topk2 != self.topk:
            idx = torch.topk(align_metric * mask_pos, self.topk2, dim=-1, largest=True).indices
            topk_idx = torch.zeros_like(mask_pos)
            topk_idx.scatter_(-1, idx, 1)
            mask_pos = mask_pos * topk_idx
            fg_mask = mask_pos.sum(-2)
        return mask_pos.argmax(-2), fg_mask, mask_pos


def dfl_loss(pred_dist, target, reg_max=16):
    """DFLoss.forward, Loss.cs:104-119.  pred_dist (n*4, reg_max) logits, target (n, 4)."""
    target = target.clamp(0, reg_max - 1 - 0.01)
    tl = target.long()
    tr = tl + 1
    wl = tr - target
    wr = 1 - wl
    return (F.cross_entropy(pred_dist, tl.view(-1), reduction="none").view(tl.shape) * wl +
            F.cross_entropy(pred_dist, tr.view(-1), reduction="none").view(tl.shape) * wr).mean(-1, keepdim=True)


def bbox_loss(pred_dist, pred_bboxes, anchor_points, target_bboxes, target_scores, target_scores_sum, fg_mask, reg_max=16):
    """BboxLoss.forward with DFL, Loss.cs:134-152."""
    weight = target_scores.sum(-1)[fg_mask].unsqueeze(-1)
    iou = bbox_iou_ciou(pred_bboxes[fg_mask], target_bboxes[fg_mask])
    loss_iou = ((1.0 - iou) * weight).sum() / target_scores_sum
    target_ltrb = bbox2dist(anchor_points, target_bboxes, reg_max - 1)
    loss_dfl = dfl_loss(pred_dist[fg_mask].view(-1, reg_max), target_ltrb[fg_mask], reg_max) * weight
    return loss_iou, loss_dfl.sum() / target_scores_sum


class V8DetectionLoss:
    """v8DetectionLoss, Loss.cs:328-485.  `preds` = {"boxes": (b, 4*reg_max, A) raw distribution logits,
    "scores": (b, nc, A) class logits, "feats": the three head feature maps (for the anchor grid)};
    `batch` = {"batch_idx": (n,), "cls": (n,), "bboxes": (n, 4) normalised xywh}.  Returns (loss * batch, items)."""

    def __init__(self, nc, reg_max=16, stride=(8, 16, 32), tal_topk=10, hyp_box=7.5, hyp_cls=0.5, hyp_dfl=1.5):
        self.nc, self.reg_max, self.stride = nc, reg_max, list(stride)
        self.hyp = (hyp_box, hyp_cls, hyp_dfl)
        self.assigner = TaskAlignedAssigner(topk=tal_topk, num_classes=nc, alpha=0.5, beta=6.0, stride=self.stride)
        self.proj = torch.arange(reg_max, dtype=torch.float32)

    def preprocess(self, targets, batch_size, scale_tensor):
        """Loss.cs:363-389: (n, 6) rows [img, cls, xywh] -> (b, max_n, 5) [cls, xyxy pixels], zero padded."""
        nl, ne = targets.shape
        if nl == 0:
            return torch.zeros(batch_size, 0, ne - 1)
        batch_idx = targets[:, 0].long()
        counts = batch_idx.unique(return_counts=True)[1]
        out = torch.zeros(batch_size, int(counts.max()), ne - 1)
        offsets = torch.zeros(batch_size + 1, dtype=torch.int64).scatter_add_(0, batch_idx + 1, torch.ones_like(batch_idx)).cumsum(0)
        within = torch.arange(nl) - offsets[batch_idx]
        out[batch_idx, within] = targets[:, 1:]
        out[..., 1:5] = xywh2xyxy(out[..., 1:5] * scale_tensor)
        return out

    def bbox_decode(self, anchor_points, pred_dist):
        """Loss.cs:397-408."""
        b, a, c = pred_dist.shape
        pred_dist = pred_dist.view(b, a, 4, c // 4).softmax(3).matmul(self.proj.type(pred_dist.dtype))
        return dist2bbox(pred_dist, anchor_points, xywh=False)

    def assign(self, preds, batch):
        """Target side of get_assigned_targets_and_loss (Loss.cs:410-441): no gradient flows through it."""
        pred_distri = preds["boxes"].permute(0, 2, 1).contiguous()
        pred_scores = preds["scores"].permute(0, 2, 1).contiguous()
        anchor_points, stride_tensor = make_anchors(preds["feats"], self.stride, 0.5)
        batch_size = pred_scores.shape[0]
        imgsz = torch.tensor(preds["feats"][0].shape[2:], dtype=pred_scores.dtype) * self.stride[0]
        targets = torch.cat((batch["batch_idx"].view(-1, 1), batch["cls"].view(-1, 1), batch["bboxes"]), 1)
        targets = self.preprocess(targets.float(), batch_size, imgsz[[1, 0, 1, 0]].float()).to(pred_scores.dtype)
        gt_labels, gt_bboxes = targets.split((1, 4), 2)
        mask_gt = gt_bboxes.sum(2, keepdim=True).gt(0.0).to(gt_bboxes.dtype)
        pred_bboxes = self.bbox_decode(anchor_points, pred_distri)
        _, target_bboxes, target_scores, fg_mask, target_gt_idx = self.assigner.forward(
            pred_scores.detach().sigmoid(), (pred_bboxes.detach() * stride_tensor).type(gt_bboxes.dtype),
            anchor_points * stride_tensor, gt_labels, gt_bboxes, mask_gt)
        return fg_mask.bool(), target_gt_idx, target_bboxes, target_scores

    def loss_from_targets(self, preds, targets):
        """Differentiable side (Loss.cs:443-465) for a FIXED assignment: (box, cls, dfl) * gains."""
        fg_mask, _, target_bboxes, target_scores = targets
        loss = torch.zeros(3, dtype=preds["scores"].dtype)  # float32 in the reference; follows the inputs so that
        pred_distri = preds["boxes"].permute(0, 2, 1).contiguous()  # double-precision gradient checks are possible
        pred_scores = preds["scores"].permute(0, 2, 1).contiguous()
        anchor_points, stride_tensor = make_anchors(preds["feats"], self.stride, 0.5)
        pred_bboxes = self.bbox_decode(anchor_points, pred_distri)
        target_scores_sum = max(float(target_scores.sum()), 1.0)
        loss[1] = F.binary_cross_entropy_with_logits(pred_scores, target_scores.to(pred_scores.dtype), reduction="none").sum() / target_scores_sum
        if fg_mask.sum() > 0:
            loss[0], loss[2] = bbox_loss(pred_distri, pred_bboxes, anchor_points, target_bboxes / stride_tensor, target_scores,
                                         target_scores_sum, fg_mask, self.reg_max)
        return loss * torch.tensor(self.hyp, dtype=loss.dtype)

    def assigned_targets_and_loss(self, preds, batch):
        """get_assigned_targets_and_loss, Loss.cs:410-466."""
        targets = self.assign(preds, batch)
        return targets, self.loss_from_targets(preds, targets)

    def __call__(self, preds, batch):
        """Loss.cs:468-483: (loss * batch_size, loss.detach())."""
        _, loss = self.assigned_targets_and_loss(preds, batch)
        return loss * preds["boxes"].shape[0], loss.detach()


# ---------------------------------------------------------------------------------------------------------------------
# v8SegmentationLoss: the instance-mask term (Loss.cs:688-865).  The box / cls / dfl terms are v8DetectionLoss's
# (get_assigned_targets_and_loss, :411-468); what the segmentation criterion adds is calculate_segmentation_loss
# (:806-861) with single_mask_loss (:787-795), driven by the assigner's fg_mask / target_gt_idx / target_bboxes.
# ---------------------------------------------------------------------------------------------------------------------
def single_mask_loss(gt_mask, pred, proto, xyxy, area):
    """Loss.cs:787-795.  gt_mask (n, H, W), pred (n, nm), proto (nm, H, W), xyxy (n, 4) in mask pixels, area (n,)."""
    from .ops import crop_mask  # the CUDA / n >= 50 branch of Ops.crop_mask (Ops.cs:437-447)
    pred_mask = torch.einsum("in,nhw->ihw", pred, proto)
    loss = F.binary_cross_entropy_with_logits(pred_mask, gt_mask, reduction="none")
    return (crop_mask(loss, xyxy).mean(dim=(1, 2)) / area).sum()


def calculate_segmentation_loss(fg_mask, masks, target_gt_idx, target_bboxes, proto, pred_masks, imgsz):
    """Loss.cs:806-861, overlap_mask = true (the reference's default, :694): masks (B, H, W) holds instance index + 1.
    imgsz = (H_img, W_img).  -> scalar: sum over images of single_mask_loss / number of foreground anchors."""
    mask_h, mask_w = proto.shape[2], proto.shape[3]
    loss = torch.zeros(1)
    wh = torch.stack((imgsz[1], imgsz[0], imgsz[1], imgsz[0]))
    tbn = target_bboxes / wh
    marea = xyxy2xywh(tbn)[..., 2:].prod(2)
    mxyxy = tbn * torch.tensor([mask_w, mask_h, mask_w, mask_h], dtype=tbn.dtype)
    for i in range(fg_mask.shape[0]):
        if bool(fg_mask[i].any()):
            mask_idx = target_gt_idx[i][fg_mask[i]]
            gt_mask = (masks[i] == (mask_idx + 1).view(-1, 1, 1)).float()
            loss = loss + single_mask_loss(gt_mask, pred_masks[i][fg_mask[i]], proto[i], mxyxy[i][fg_mask[i]], marea[i][fg_mask[i]])
        else:
            loss = loss + (proto * 0).sum() + (pred_masks * 0).sum()
    return loss.sum() / fg_mask.sum()


def segmentation_mask_loss(fg_mask, target_gt_idx, target_bboxes, masks, proto, mask_coefficient, imgsz, hyp_box=7.5):
    """loss[1] of v8SegmentationLoss.loss (Loss.cs:712-786): mask_coefficient (B, nm, A) as the head returns it, proto
    (B, nm, H, W) at the resolution of `masks`.  -> (loss[1] * batch_size, loss[1]) as the criterion returns them."""
    pred_masks = mask_coefficient.permute(0, 2, 1).contiguous()
    B = proto.shape[0]
    if float(fg_mask.sum()) > 0:
        assert masks.shape[-2:] == proto.shape[-2:], "the interpolate branch (:739-743) is not restated"
        item = calculate_segmentation_loss(fg_mask, masks.float(), target_gt_idx, target_bboxes, proto, pred_masks, imgsz)
    else:
        item = (proto * 0).sum() + (pred_masks * 0).sum()
    item = item * hyp_box
    return item * B, item.detach()
