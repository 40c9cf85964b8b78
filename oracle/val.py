"""Oracle restatement of the validation matching (test infrastructure only).

box_iou            Utils/Metrics.cs:16-34
match_predictions  Models/YoloBaseTaskModel.cs:377-446 (incl. GetUniqueMatches / GetUniqueByColumn: first occurrence per
                   unique value, rows returned in the order of the sorted unique values)
"""
import torch


def box_iou(box1, box2, eps=1e-7):
    a1, a2 = box1.float().unsqueeze(1).chunk(2, 2)
    b1, b2 = box2.float().unsqueeze(0).chunk(2, 2)
    inter = (torch.min(a2, b2) - torch.max(a1, b1)).clamp_(0).prod(2)
    return inter / ((a2 - a1).prod(2) + (b2 - b1).prod(2) - inter + eps)


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
