"""OBB / Pose decode tails and rotated NMS: oracle restatements (oracle/heads.py) against direct scalar definitions on
the CPU, and the GPU kernels (csrc/heads.cu) against the oracle.  Transcendentals differ by a few ulps between CUDA
and libm, so values are compared at 1e-5 and the NMS cases keep every pairwise IoU away from the threshold."""
import math

import pytest
import torch

from oracle import heads as oh


def _anchors(levels=((8, 8, 8.0), (4, 4, 16.0), (2, 2, 32.0))):
    pts, st = [], []
    for h, w, s in levels:
        sy, sx = torch.meshgrid(torch.arange(h) + 0.5, torch.arange(w) + 0.5, indexing="ij")
        pts.append(torch.stack([sx.reshape(-1), sy.reshape(-1)], 0))
        st.append(torch.full((h * w,), s))
    return torch.cat(pts, 1).float(), torch.cat(st).float()


def _obbs(n, seed, spread=400.0):
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand(n, 2, generator=g) * spread
    wh = torch.rand(n, 2, generator=g) * 60 + 10
    r = (torch.rand(n, 1, generator=g) - 0.25) * math.pi
    return torch.cat([xy, wh, r], 1)


def test_oracle_obb_and_pose_decode_definitions():
    anchors, strides = _anchors()
    A = anchors.shape[1]
    g = torch.Generator().manual_seed(0)
    box, cls, ang = torch.randn(2, 64, A, generator=g), torch.randn(2, 5, A, generator=g), torch.randn(2, 1, A, generator=g)
    y = oh.obb_inference(box, cls, ang, anchors, strides)
    assert y.shape == (2, 4 + 5 + 1, A)
    # scalar definition at one anchor
    b, a = 1, 37
    d = [sum(j * p for j, p in enumerate(torch.softmax(box[b, s * 16:(s + 1) * 16, a], 0).tolist())) for s in range(4)]
    th = (1 / (1 + math.exp(-float(ang[b, 0, a]))) - 0.25) * math.pi
    xf, yf = (d[2] - d[0]) / 2, (d[3] - d[1]) / 2
    ref = [(xf * math.cos(th) - yf * math.sin(th) + float(anchors[0, a])) * float(strides[a]),
           (xf * math.sin(th) + yf * math.cos(th) + float(anchors[1, a])) * float(strides[a]),
           (d[0] + d[2]) * float(strides[a]), (d[1] + d[3]) * float(strides[a])]
    torch.testing.assert_close(y[b, :4, a], torch.tensor(ref), rtol=1e-5, atol=1e-4)
    assert abs(float(y[b, 9, a]) - th) < 1e-6
    k = torch.randn(2, 17 * 3, A, generator=g)
    z = oh.kpts_decode(k, anchors, strides, 3)
    assert abs(float(z[0, 3 * 4, 5]) - (float(k[0, 12, 5]) * 2 + float(anchors[0, 5]) - 0.5) * float(strides[5])) < 1e-5
    assert abs(float(z[0, 3 * 4 + 2, 5]) - 1 / (1 + math.exp(-float(k[0, 14, 5])))) < 1e-6


def test_oracle_probiou_and_rotated_nms_properties():
    o = _obbs(40, 1)
    iou = oh.batch_probiou(o, o)
    assert iou.shape == (40, 40) and bool((iou.diagonal() > 0.99).all()) and bool((iou <= 1.0 + 1e-6).all()) and bool((iou >= 0).all())
    torch.testing.assert_close(iou, iou.T, rtol=1e-4, atol=1e-5)
    s = torch.rand(40, generator=torch.Generator().manual_seed(2))
    keep = oh.nms_rotated(o, s, 0.3)
    order = torch.argsort(s, descending=True)
    # direct definition: box j (in score order) survives iff no higher-scored box overlaps it at >= thr
    ref = [int(order[j]) for j in range(40) if not any(float(iou[order[i], order[j]]) >= 0.3 for i in range(j))]
    assert keep.tolist() == ref


@pytest.mark.gpu
def test_gpu_obb_pose_decode_vs_oracle():
    import yolosharp_b200.engine as E
    anchors, strides = _anchors(((80, 80, 8.0), (40, 40, 16.0), (20, 20, 32.0)))
    A = anchors.shape[1]
    g = torch.Generator().manual_seed(3)
    box, cls, ang = torch.randn(3, 64, A, generator=g) * 2, torch.randn(3, 15, A, generator=g), torch.randn(3, 1, A, generator=g)
    ref = oh.obb_inference(box, cls, ang, anchors, strides)
    out = E.obb_decode(box.cuda(), cls.cuda(), ang.cuda(), anchors.cuda(), strides.cuda()).cpu()
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=2e-4)
    for ndim in (3, 2):
        k = torch.randn(2, 17 * ndim, A, generator=g)
        torch.testing.assert_close(E.pose_decode(k.cuda(), anchors.cuda(), strides.cuda(), ndim).cpu(),
                                   oh.kpts_decode(k, anchors, strides, ndim), rtol=1e-6, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("n,thr,seed", [(300, 0.45, 4), (2000, 0.3, 5), (1, 0.45, 6), (0, 0.45, 7), (5000, 0.6, 8)])
def test_gpu_probiou_and_rotated_nms_vs_oracle(n, thr, seed):
    import yolosharp_b200.engine as E
    o = _obbs(n, seed, spread=400.0 if n < 3000 else 1500.0)
    s = torch.rand(n, generator=torch.Generator().manual_seed(seed + 100))
    keep = E.nms_rotated(o.cuda(), s.cuda(), thr).cpu()
    if n == 0:
        assert keep.numel() == 0
        return
    iou = oh.batch_probiou(o, o)
    gi = E.probiou(o.cuda(), o.cuda()).cpu()
    torch.testing.assert_close(gi, iou, rtol=1e-4, atol=2e-5)
    if bool(((iou - thr).abs() < 1e-4).any()):
        # some of the n^2 pairs sit within CUDA-vs-libm rounding of the threshold: check the suppression LOGIC on the
        # IoU matrix the GPU itself computes (same device function), values are pinned by the comparison above
        order = torch.argsort(s, descending=True)
        m = gi[order][:, order].triu(1) >= thr
        ref = order[(m.sum(0) <= 0).nonzero().squeeze(-1)]
    else:
        ref = oh.nms_rotated(o, s, thr)
    assert keep.tolist() == ref.tolist()
