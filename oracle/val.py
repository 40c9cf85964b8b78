"""Oracle restatement of the validation matching (test infrastructure only).

box_iou            Utils/Metrics.cs:16-34
mask_iou           Utils/Metrics.cs:120-125
match_predictions  Models/YoloBaseTaskModel.cs:377-446 (incl. GetUniqueMatches / GetUniqueByColumn: first occurrence per
                   unique value, rows returned in the order of the sorted unique values)
ap_per_class       Utils/Metrics.cs:308-384, with compute_ap :395-421, interp :424-468, smooth :475-487.
                   The reference calls torch.argsort (unstable) in three places; wherever equal keys occur (equal
                   confidences, recall plateaus in compute_ap's mrec) its result depends on that unspecified order.  This
                   restatement - and the CUDA kernel it checks - use the STABLE order (ties keep their input order).
"""
import torch


def box_iou(box1, box2, eps=1e-7):
    a1, a2 = box1.float().unsqueeze(1).chunk(2, 2)
    b1, b2 = box2.float().unsqueeze(0).chunk(2, 2)
    inter = (torch.min(a2, b2) - torch.max(a1, b1)).clamp_(0).prod(2)
    return inter / ((a2 - a1).prod(2) + (b2 - b1).prod(2) - inter + eps)


def mask_iou(mask1, mask2, eps=1e-7):
    """Utils/Metrics.cs:120-125: mask1 (N, n), mask2 (M, n) -> (N, M)."""
    inter = torch.matmul(mask1, mask2.T).clamp_(0)
    union = (mask1.sum(1)[..., None] + mask2.sum(1)[None]) - inter
    return inter / (union + eps)


def _unique_by_column(matches, col):
    vals = matches[:, col]
    uniq, inv = vals.unique(return_inverse=True)
    first = torch.full((uniq.shape[0],), -1, dtype=torch.long)
    for i in range(vals.shape[0]):
        if first[inv[i]] == -1:
            first[inv[i]] = i
    return matches.index_select(0, first)


def match_predictions(pred_classes, true_classes, iou):
    iouv = torch.linspace(0.5, 0.95, 10, dtype=torch.float32)
    correct = torch.zeros((pred_classes.shape[0], iouv.shape[0]), dtype=torch.bool)
    correct_class = true_classes[:, None] == pred_classes
    iou = iou * correct_class
    for i in range(iouv.numel()):
        matches = torch.nonzero(iou >= float(iouv[i]))
        if matches.shape[0] > 0:
            if matches.shape[0] > 1:
                matches = matches[iou[matches[:, 0], matches[:, 1]].argsort(descending=True)]
                matches = _unique_by_column(_unique_by_column(matches, 1), 0)
            correct[matches[:, 1], i] = True
    return correct


def interp(x, xp, fp, left=0.0):
    """Metrics.cs:424-468: np.interp-like, but `left` is a constant (default 0, NOT fp[0]) and the right side is fp[-1]."""
    idx = torch.argsort(xp, stable=True)
    xps, fps = xp[idx].contiguous(), fp[idx].contiguous()
    res = torch.empty_like(x)
    res[x >= xps[-1]] = fps[-1]
    res[x <= xps[0]] = left
    interior = (x > xps[0]) & (x < xps[-1])
    if int(interior.sum()) > 0:
        xi = x[interior]
        k = (torch.searchsorted(xps, xi) - 1).clamp(0, xps.shape[0] - 2)
        x0, x1, y0, y1 = xps[k], xps[k + 1], fps[k], fps[k + 1]
        t = (xi - x0) / (x1 - x0)
        res[interior] = y0 + t * (y1 - y0)
    return res


def compute_ap(recall, precision):
    """Metrics.cs:395-421 ("interp" method: 101-point COCO interpolation, trapezoid)."""
    mrec = torch.cat((torch.tensor([0.0]), recall, torch.tensor([1.0])))
    mpre = torch.cat((torch.tensor([1.0]), precision, torch.tensor([0.0])))
    mpre = mpre.flip(0).cummax(0).values.flip(0)
    x = torch.linspace(0, 1, 101)
    return float(torch.trapezoid(interp(x, mrec, mpre), x)), mpre, mrec


def smooth(y, f=0.05):
    """Metrics.cs:475-487: box filter; BOTH pads repeat y[0] (the reference's `ones * y[0]` is used on either side)."""
    nf = int(y.shape[0] * f * 2) // 2 * 2 + 1
    p = torch.ones(nf // 2) * y[0]
    yp = torch.cat((p, y, p))
    return torch.nn.functional.conv1d(yp.view(1, 1, -1), (torch.ones(nf) / nf).view(1, 1, -1)).view(-1)


def ap_per_class(tp, conf, pred_cls, target_cls, eps=1e-16):
    """Metrics.cs:308-384.  tp (n, T) bool, conf (n,), pred_cls (n,), target_cls (m,) -> dict of the reference's outputs."""
    tp, conf, pred_cls = tp.bool(), conf.float(), pred_cls.float()
    ii = torch.argsort(-conf, stable=True)
    tp, conf, pred_cls = tp[ii], conf[ii], pred_cls[ii]
    unique_classes, nt = torch.unique(target_cls.float(), return_counts=True)
    nc, T = unique_classes.shape[0], tp.shape[1]
    x = torch.linspace(0, 1, 1000)
    prec_values = []
    ap = torch.zeros((nc, T))
    p_curve, r_curve = torch.zeros((nc, 1000)), torch.zeros((nc, 1000))
    for ci in range(nc):
        i = pred_cls == unique_classes[ci]
        n_l, n_p = int(nt[ci]), int(i.sum())
        if n_p == 0 or n_l == 0:
            continue
        fpc = (~tp[i]).cumsum(0)
        tpc = tp[i].cumsum(0)
        recall = tpc / (n_l + eps)
        r_curve[ci] = interp(-x, -conf[i], recall[:, 0], left=0)
        precision = tpc / (tpc + fpc)
        p_curve[ci] = interp(-x, -conf[i], precision[:, 0], left=1)
        for j in range(T):
            ap[ci, j], mpre, mrec = compute_ap(recall[:, j], precision[:, j])
            if j == 0:
                prec_values.append(interp(x, mrec, mpre))
    if not prec_values:
        prec_values = [torch.zeros(1000)]
    f1_curve = 2 * p_curve * r_curve / (p_curve + r_curve + eps)
    best = int(smooth(f1_curve.mean(0), 0.1).argmax())
    p, r, f1 = p_curve[:, best], r_curve[:, best], f1_curve[:, best]
    tpn = (r * nt).round()
    fpn = (tpn / (p + eps) - tpn).round()
    return {"tp": tpn, "fp": fpn, "p": p, "r": r, "f1": f1, "ap": ap, "unique_classes": unique_classes.int(), "p_curve": p_curve,
            "r_curve": r_curve, "f1_curve": f1_curve, "x": x, "prec_values": torch.stack(prec_values), "best": best}
