"""Oracle restatement of Utils/Ops.cs post-processing (test infrastructure only).

``non_max_suppression`` follows Utils/Ops.cs:239-371 line by line; the greedy
suppress core is ``torchvision.ops.nms`` exactly as the reference calls it
(Ops.cs:357; TorchVision 0.105.2 there, torchvision 0.26 CPU here - same
contract: score-descending stable order, suppress when IoU > thr,
IoU = inter / (a1 + a2 - inter), no epsilon).  ``greedy_nms_numpy`` is an
independent scalar restatement of that contract used to cross-check it.
"""
import numpy as np
import torch
import torchvision


def xywh2xyxy(x):
    """Ops.cs:68-81."""
    assert x.shape[-1] == 4
    y = torch.zeros_like(x)
    y[..., 0] = x[..., 0] - x[..., 2] / 2
    y[..., 1] = x[..., 1] - x[..., 3] / 2
    y[..., 2] = x[..., 0] + x[..., 2] / 2
    y[..., 3] = x[..., 1] + x[..., 3] / 2
    return y


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, agnostic=False, max_det=300,
                        nc=0, max_nms=30000, max_wh=7680):
    """Ops.cs:239-371 (non-rotated, non-end2end path).  Returns (output, keepi):
    per image a (n,6+extra) tensor [x1,y1,x2,y2,conf,cls,extra...] and the kept
    anchor indices.  `agnostic` is accepted and ignored, as in the reference
    (the offset `c` is always applied, Ops.cs:345)."""
    if conf_thres < 0 or conf_thres > 1:
        raise ValueError(f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0")
    if iou_thres < 0 or iou_thres > 1:
        raise ValueError(f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0")
    prediction = prediction.clone()  # the reference converts in place; keep the caller's tensor
    bs = prediction.shape[0]
    nc = nc or prediction.shape[1] - 4
    extra = prediction.shape[1] - nc - 4
    mi = 4 + nc
    xc = prediction[:, 4:mi].amax(1) > conf_thres
    xinds = torch.stack([torch.arange(xc.shape[1]) for _ in range(bs)], 0).unsqueeze(-1)
    prediction = prediction.transpose(-1, -2)
    prediction[..., :4] = xywh2xyxy(prediction[..., :4])
    output = [torch.zeros((0, 6 + extra)) for _ in range(bs)]
    keepi = [torch.zeros((0,), dtype=torch.long) for _ in range(bs)]
    for xi in range(bs):
        x = prediction[xi]
        xk = xinds[xi]
        filt = xc[xi]
        x, xk = x[filt], xk[filt]
        if x.shape[0] == 0:
            continue
        box, cls, mask = x.split((4, nc, extra), 1)
        conf, j = cls.max(1, keepdim=True)
        filt = conf.view(-1) > conf_thres
        x = torch.cat((box, conf, j.float(), mask), 1)[filt]
        xk = xk[filt]
        n = x.shape[0]
        if n == 0:
            continue
        if n > max_nms:
            filt = x[:, 4].argsort(descending=True, stable=True)[:max_nms]
            x, xk = x[filt], xk[filt]
        c = x[:, 5:6] * max_wh
        scores = x[:, 4]
        boxes = x[:, :4] + c
        # The reference passes a C# `float` (Ops.cs:241) that is widened to the kernel's double
        # threshold; a Python float would compare against 0.3 instead of (double)0.3f, which
        # differs exactly when an IoU equals the fp32 threshold.
        i = torchvision.ops.nms(boxes, scores, float(np.float32(iou_thres)))
        i = i[:max_det]
        output[xi], keepi[xi] = x[i], xk[i].reshape(-1)
    return output, keepi


def greedy_nms_numpy(boxes, scores, iou_thres):
    """Independent scalar restatement of torchvision's CPU nms kernel contract
    (fp32 arithmetic, stable score-descending order, suppress on ovr > thr)."""
    boxes = np.asarray(boxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    order = np.argsort(-scores, kind="stable")
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(len(boxes), dtype=bool)
    keep = []
    thr = np.float32(iou_thres)
    for _i in range(len(order)):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr > thr]] = True
    return np.asarray(keep, dtype=np.int64)


def crop_mask(masks, boxes):
    """Ops.cs:409-451, CUDA/large-n branch (:437-447) forced: the GPU engine replaces the
    is_cuda path, and the two branches are not equivalent."""
    n, h, w = masks.shape
    x1, y1, x2, y2 = torch.chunk(boxes[:, :, None], 4, 1)
    r = torch.arange(w, dtype=x1.dtype)[None, None, :]
    c = torch.arange(h, dtype=x1.dtype)[None, :, None]
    return masks * ((r >= x1) * (r < x2) * (c >= y1) * (c < y2))


def process_mask(protos, masks_in, bboxes, shape, upsample=False):
    """Ops.cs:462-489."""
    c, mh, mw = protos.shape
    ih, iw = shape
    masks = masks_in.matmul(protos.float().view(c, -1)).view(-1, mh, mw)
    width_ratio = np.float32(mw) / np.float32(iw)
    height_ratio = np.float32(mh) / np.float32(ih)
    d = bboxes.clone()
    d[..., 0] *= float(width_ratio)
    d[..., 2] *= float(width_ratio)
    d[..., 3] *= float(height_ratio)
    d[..., 1] *= float(height_ratio)
    masks = crop_mask(masks, d)
    if upsample:
        masks = torch.nn.functional.interpolate(masks[None], size=tuple(shape), mode="bilinear", align_corners=False)[0]
    return masks.gt_(0.0)


def preprocess(img_u8_chw):
    """Models/Detector.cs:31-41: uint8 CHW RGB -> float, pad right/bottom to a multiple
    of 32 with 114, /255, add batch dim."""
    x = img_u8_chw.float().unsqueeze(0)
    h, w = x.shape[2], x.shape[3]
    ph = (32 - h % 32) % 32
    pw = (32 - w % 32) % 32
    x = torch.nn.functional.pad(x, (0, pw, 0, ph), mode="constant", value=114.0) / 255.0
    return x


def to_yolo_results(rows):
    """Models/Detector.cs:50-69: truncating int conversion into YoloResult fields."""
    res = []
    for r in rows.tolist():
        x, y = int(r[0]), int(r[1])
        rw, rh = int(r[2]) - x, int(r[3]) - y
        # C# integer division truncates toward zero
        res.append(dict(ClassID=int(r[5]), Score=float(np.float32(r[4])),
                        CenterX=x + int(rw / 2), CenterY=y + int(rh / 2), Width=rw, Height=rh))
    return res


def clip_boxes(x, shape):
    """Ops.cs:150-158: x to [0, shape[1]], y to [0, shape[0]]."""
    box = torch.zeros_like(x)
    box[..., 0] = x[..., 0].clamp(0, shape[1])
    box[..., 1] = x[..., 1].clamp(0, shape[0])
    box[..., 2] = x[..., 2].clamp(0, shape[1])
    box[..., 3] = x[..., 3].clamp(0, shape[0])
    return box


def detector_predict(model, img_u8_chw, conf, iou, pre=preprocess):
    """Models/Detector.cs:27-72 on an oracle model: -> (rows (n,6), keep, YoloResult dicts)."""
    with torch.no_grad():
        pred = model(pre(img_u8_chw))[0]["boxes"]
    out, keep = non_max_suppression(pred, conf, iou)
    return out[0], keep[0], to_yolo_results(out[0])


def segmenter_predict(model, img_u8_chw, conf, iou, nc=80, pre=preprocess):
    """Models/Segmenter.cs:28-84 on an oracle segment model: -> (rows (n,38) with boxes clipped to the original
    image, masks uint8 (n, h, w) at the ORIGINAL size, YoloResult dicts).  Masks are computed on the padded input
    and then resized - not cropped - to the original size (reference behaviour, :56-57); the `.byte()` cast
    truncates the bilinear values, so only exact ones survive."""
    h, w = int(img_u8_chw.shape[1]), int(img_u8_chw.shape[2])
    x = pre(img_u8_chw)
    with torch.no_grad():
        inf = model(x)[0]
    out, _ = non_max_suppression(inf["boxes"], conf, iou, nc=nc)
    rows = out[0].clone()
    if rows.shape[0] == 0:
        return rows, torch.zeros((0, h, w), dtype=torch.uint8), []
    masks = process_mask(inf["proto"][0], rows[:, 6:], rows[:, :4], (x.shape[2], x.shape[3]), upsample=True)
    rows[:, :4] = clip_boxes(rows[:, :4], (h, w))
    if (x.shape[2], x.shape[3]) != (h, w):
        masks = torch.nn.functional.interpolate(masks[None].float(), size=(h, w), mode="bilinear", align_corners=False)[0]
    return rows, masks.to(torch.uint8), to_yolo_results(rows)


def e2e_postprocess(y, max_det, nc, agnostic=False):
    """Head.cs:117-127 (`postprocess`) + 175-196 (`get_topk_index`): y (B, 4 + nc, A), the end2end head's decoded output
    -> (rows (B, k, 6) [box, score, class], idx (B, k) anchor of every row).  `forward` calls it as
    postprocess(y.permute(0, 2, 1)) (Head.cs:107-111)."""
    preds = y.permute(0, 2, 1)
    boxes, scores = preds.split([4, nc], dim=-1)                       # Head.cs:120-122
    batch_size, anchors, _ = scores.shape
    k = min(max_det, anchors)                                           # Head.cs:182
    if agnostic:                                                        # Head.cs:183-189
        scores, labels = scores.max(dim=-1, keepdim=True)
        scores, indices = scores.topk(k, dim=1)
        labels = labels.gather(1, indices)
        conf, idx = labels.float(), indices
    else:
        ori_index = scores.max(dim=-1).values.topk(k).indices.unsqueeze(-1)          # Head.cs:190
        scores = scores.gather(dim=1, index=ori_index.repeat(1, 1, nc))              # Head.cs:191
        scores, index = scores.flatten(1).topk(k)                                    # Head.cs:192
        idx = ori_index[torch.arange(batch_size)[..., None], torch.div(index, nc, rounding_mode="floor")]  # Head.cs:193
        scores, conf = scores[..., None], (index % nc)[..., None].float()            # Head.cs:194
    boxes = boxes.gather(dim=1, index=idx.repeat(1, 1, 4))                           # Head.cs:125
    return torch.cat([boxes, scores, conf], dim=-1), idx.squeeze(-1)                 # Head.cs:126
