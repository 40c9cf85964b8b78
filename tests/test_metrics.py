"""ap_per_class (Utils/Metrics.cs:308-384): the oracle restatement against independent formulations on the CPU, and the CUDA
entry point yb_ap_per_class against the oracle on the GPU."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import val as oval


def _case(n, m, ncls, T=10, seed=0, quant=None, extra_pred_classes=0):
    """Detections / labels with a consistent TP matrix: per class and threshold at most n_l true positives, nested over
    the thresholds (what match_predictions produces).  quant: round confidences to 1/quant (ties)."""
    g = torch.Generator().manual_seed(seed)
    conf = torch.rand(n, generator=g)
    if quant:
        conf = (conf * quant).round() / quant
    pred_cls = torch.randint(0, ncls + extra_pred_classes, (n,), generator=g)
    target_cls = torch.randint(0, ncls, (m,), generator=g)
    if ncls > 2:
        target_cls[target_cls == 1] = 0  # a class with predictions but no labels
        pred_cls[pred_cls == 2] = 3      # a class with labels but no predictions
    tp = torch.zeros((n, T), dtype=torch.bool)
    for c in range(ncls):
        idx = torch.nonzero(pred_cls == c).flatten()
        n_l = int((target_cls == c).sum())
        if idx.numel() == 0 or n_l == 0:
            continue
        # confident detections are more often right
        score = conf[idx] + 0.5 * torch.rand(idx.numel(), generator=g)
        order = idx[score.argsort(descending=True)]
        k0 = min(n_l, int(0.7 * idx.numel()))
        for j in range(T):
            kj = int(k0 * (1 - 0.08 * j))
            tp[order[:kj], j] = True
    return tp, conf, pred_cls, target_cls


def test_linspace_matches_torch():
    from yolosharp_b200 import _lib as L
    for steps in (1000, 101, 10):
        out = torch.empty(steps, dtype=torch.float32)
        L.check(L.lib().yb_linspace01(steps, C.c_void_p(out.data_ptr())))
        assert torch.equal(out, torch.linspace(0, 1, steps))


def test_oracle_interp_is_numpy_interp_with_constant_left():
    g = torch.Generator().manual_seed(1)
    xp = torch.rand(50, generator=g).sort().values
    fp = torch.rand(50, generator=g)
    x = torch.linspace(-0.2, 1.2, 400)
    got = oval.interp(x, xp, fp, left=0.25)
    want = torch.from_numpy(np.interp(x.numpy().astype(np.float64), xp.numpy().astype(np.float64), fp.numpy().astype(np.float64),
                                      left=0.25)).float()
    edge = (x <= xp[0])
    assert torch.all(got[edge] == 0.25)
    torch.testing.assert_close(got[~edge], want[~edge], rtol=1e-5, atol=1e-6)


def test_oracle_ap_known_answers():
    # every detection right, all labels found: precision 1 everywhere, recall reaches 1 -> 101-point AP = 0.99: the
    # reference's interp returns `left` = 0 at recall 0 and mpre's closing 0 at recall 1, so both end trapezoids are half
    n = 40
    tp = torch.ones((n, 10), dtype=torch.bool)
    o = oval.ap_per_class(tp, torch.linspace(0.9, 0.1, n), torch.zeros(n), torch.zeros(n))
    assert o["ap"].shape == (1, 10) and abs(float(o["ap"][0, 0]) - 0.99) < 1e-6
    assert o["tp"].tolist() == [float(n)] and o["fp"].tolist() == [0.0]
    # nothing right
    o = oval.ap_per_class(torch.zeros((n, 10), dtype=torch.bool), torch.linspace(0.9, 0.1, n), torch.zeros(n), torch.zeros(n))
    assert float(o["ap"].abs().max()) == 0.0
    # no detections of a labelled class: its rows stay zero and prec_values falls back to one zero row
    o = oval.ap_per_class(torch.zeros((3, 10), dtype=torch.bool), torch.tensor([0.5, 0.4, 0.3]), torch.tensor([5., 5., 5.]), torch.zeros(4))
    assert o["unique_classes"].tolist() == [0] and o["prec_values"].shape == (1, 1000) and float(o["p_curve"].abs().max()) == 0.0


def test_oracle_ap_against_direct_loop():
    """AP of one class / threshold recomputed with plain Python from the definitions (envelope, 101 points, trapezoid)."""
    tp, conf, pc, tc = _case(300, 80, 4, seed=3)
    o = oval.ap_per_class(tp, conf, pc, tc)
    c, j = 0, 2
    sel = (pc == c).nonzero().flatten()
    sel = sel[torch.argsort(-conf[sel], stable=True)]
    n_l = int((tc == c).sum())
    t = tp[sel, j].numpy()
    tpc = np.cumsum(t)
    rec = (tpc / n_l).astype(np.float32)
    pre = (tpc / np.arange(1, len(t) + 1)).astype(np.float32)
    mrec = np.concatenate(([0.0], rec, [1.0]))
    mpre = np.concatenate(([1.0], pre, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    xs = np.linspace(0, 1, 101)
    ys = np.interp(xs, mrec, mpre)
    ys[0] = 0.0  # the reference's interp: `left` (0) at x <= mrec[0]
    ys[-1] = mpre[-1]
    want = np.trapezoid(ys, xs)
    ci = o["unique_classes"].tolist().index(c)
    assert abs(float(o["ap"][ci, j]) - want) < 2e-5


def _cmp(o, g, atol=2e-6):
    assert g["unique_classes"].cpu().tolist() == o["unique_classes"].tolist()
    bad = []
    for k in ("ap", "p_curve", "r_curve", "f1_curve", "prec_values", "p", "r", "f1", "tp", "fp"):
        a, b = g[k].cpu(), o[k].float()
        if a.shape != b.shape:
            bad.append(f"{k}: shape {tuple(a.shape)} != {tuple(b.shape)}")
            continue
        if a.numel() and not torch.allclose(a, b, rtol=1e-5, atol=atol):
            d = (a - b).abs()
            bad.append(f"{k}: max abs diff {float(d.max()):.3e} at {tuple(int(v) for v in (d == d.max()).nonzero()[0])}, "
                       f"{int((d > atol + 1e-5 * b.abs()).sum())} of {a.numel()} off")
    if g["best"] != o["best"]:
        bad.append(f"best: {g['best']} != {o['best']}")
    assert not bad, "; ".join(bad)
    assert torch.equal(g["x"], o["x"])


@pytest.mark.gpu
@pytest.mark.parametrize("n,m,ncls,quant,extra", [(3000, 700, 12, None, 0), (5000, 900, 80, 200, 0), (40, 30, 3, None, 0),
                                                  (70000, 9000, 20, 1000, 0), (1500, 300, 6, None, 3), (1, 5, 1, None, 0)])
def test_ap_per_class_matches_oracle_gpu(n, m, ncls, quant, extra):
    from yolosharp_b200 import engine as E
    tp, conf, pc, tc = _case(n, m, ncls, seed=n, quant=quant, extra_pred_classes=extra)
    o = oval.ap_per_class(tp, conf, pc, tc)
    g = E.ap_per_class(tp.cuda(), conf.cuda(), pc.cuda(), tc.cuda(), max_classes=max(80, ncls + extra))
    _cmp(o, g)


@pytest.mark.gpu
def test_ap_per_class_empty_inputs_gpu():
    from yolosharp_b200 import engine as E
    # labels but no detections at all
    g = E.ap_per_class(torch.zeros((0, 10), dtype=torch.bool).cuda(), torch.zeros(0).cuda(), torch.zeros(0).cuda(),
                       torch.tensor([2, 2, 5]).cuda())
    assert g["unique_classes"].cpu().tolist() == [2, 5] and float(g["ap"].abs().max()) == 0.0
    assert g["prec_values"].shape == (1, 1000) and g["tp"].cpu().tolist() == [0.0, 0.0]
    # detections but no labels: no unique classes
    g = E.ap_per_class(torch.ones((4, 10), dtype=torch.bool).cuda(), torch.rand(4).cuda(), torch.zeros(4).cuda(), torch.zeros(0).cuda())
    assert g["ap"].shape == (0, 10) and g["unique_classes"].numel() == 0


def test_oracle_reproduces_golden_ap():
    """tests/golden/r2_val_seg.npz (make_golden_r2b.py): the committed ap_per_class vectors."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "r2_val_seg.npz"))
    o = oval.ap_per_class(torch.from_numpy(g["ap_tp"]), torch.from_numpy(g["ap_conf"]), torch.from_numpy(g["ap_pred_cls"]),
                          torch.from_numpy(g["ap_target_cls"]))
    assert o["unique_classes"].tolist() == g["ap_unique"].tolist() and o["best"] == int(g["ap_best"])
    for k, gk in (("ap", "ap_ap"), ("p", "ap_p"), ("r", "ap_r"), ("f1", "ap_f1"), ("tp", "ap_tpn"), ("fp", "ap_fpn"), ("p_curve", "ap_p_curve"),
                  ("r_curve", "ap_r_curve"), ("prec_values", "ap_prec_values")):
        torch.testing.assert_close(o[k].float(), torch.from_numpy(g[gk]).float(), rtol=1e-6, atol=1e-7)


def test_oracle_mask_iou_known_answers():
    a = torch.tensor([[1., 1., 0., 0.], [0., 0., 0., 0.]])
    b = torch.tensor([[1., 0., 0., 0.], [1., 1., 1., 1.], [0., 0., 1., 1.]])
    got = oval.mask_iou(a, b)
    want = torch.tensor([[0.5, 0.5, 0.0], [0.0, 0.0, 0.0]])
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="yb_mask_iou was written after this round's GPU minutes were spent: its first run on hardware is the "
                                        "driver's; a pass shows as XPASS")
def test_mask_iou_matches_oracle_gpu():
    from yolosharp_b200 import engine as E
    g = torch.Generator().manual_seed(0)
    for n1, n2, px in ((7, 12, 160 * 160), (1, 1, 33), (30, 3, 1000)):
        a = (torch.rand(n1, px, generator=g) > 0.6).float()
        b = (torch.rand(n2, px, generator=g) > 0.5).float()
        b[0] = 0  # an empty mask: 0 / eps
        got = E.mask_iou(a.cuda(), b.cuda()).cpu()
        torch.testing.assert_close(got, oval.mask_iou(a, b), rtol=1e-6, atol=1e-7)
    assert E.mask_iou(torch.zeros(0, 16).cuda(), torch.zeros(3, 16).cuda()).shape == (0, 3)
