"""End2end post-processing (Head.cs:117-127, 175-196): oracle restatement vs a scalar restatement on the CPU, and the GPU
kernel (csrc/topk.cu, yb_topk_postprocess) vs the oracle.  Scores are distinct in every case that compares rows one to
one (torch.topk leaves ties unspecified); a separate case checks that ties produce a valid selection."""
import pytest
import torch

from oracle import ops as oops


def _pred(B, nc, A, seed, ties=False):
    g = torch.Generator().manual_seed(seed)
    box = torch.rand(B, 4, A, generator=g) * 640
    if ties:
        sc = torch.randint(0, 8, (B, nc, A), generator=g).float() / 8
    else:
        n = nc * A  # a random permutation of n equally spaced values per image: distinct fp32 scores (n < 2^23)
        sc = torch.stack([((torch.randperm(n, generator=g).double() + 0.5) / n).float().view(nc, A) for _ in range(B)])
    return torch.cat([box, sc], 1).contiguous()


def _scalar(y, max_det, nc):
    """Direct definition: top-k anchors by best score, then top-k (anchor, class) pairs among them."""
    B, _, A = y.shape
    k = min(max_det, A)
    rows = []
    for b in range(B):
        s = y[b, 4:4 + nc]                                  # (nc, A)
        best = s.max(0).values
        anchors = sorted(range(A), key=lambda a: -float(best[a]))[:k]
        pairs = sorted(((float(s[c, a]), a, c) for a in anchors for c in range(nc)), key=lambda t: -t[0])[:k]
        rows.append(torch.tensor([[*y[b, :4, a].tolist(), v, float(c)] for v, a, c in pairs]))
    return torch.stack(rows)


def test_oracle_e2e_postprocess_vs_scalar():
    y = _pred(2, 5, 40, 0)
    out, idx = oops.e2e_postprocess(y, 12, 5)
    ref = _scalar(y, 12, 5)
    assert out.shape == (2, 12, 6) and idx.shape == (2, 12)
    torch.testing.assert_close(out, ref, rtol=0, atol=1e-6)
    # max_det larger than the number of anchors: k = A
    out2, _ = oops.e2e_postprocess(y, 300, 5)
    assert out2.shape == (2, 40, 6)
    # the agnostic branch: one row per selected anchor, its best class
    outa, idxa = oops.e2e_postprocess(y, 7, 5, agnostic=True)
    best, cls = y[:, 4:9].max(1)
    top = best.topk(7).indices
    assert torch.equal(idxa, top) and torch.equal(outa[..., 5], cls.gather(1, top).float())


@pytest.mark.gpu
@pytest.mark.parametrize("B,nc,A,max_det", [(3, 80, 8400, 300), (2, 80, 2100, 300), (1, 3, 100, 300), (2, 1, 8400, 100),
                                            (1, 80, 33600, 300), (2, 7, 777, 1000)])
def test_gpu_topk_postprocess_vs_oracle(B, nc, A, max_det):
    import yolosharp_b200.engine as E
    y = _pred(B, nc, A, 1)
    ref, ridx = oops.e2e_postprocess(y, max_det, nc)
    out, idx = E.topk_postprocess(y.cuda(), max_det, nc)
    assert torch.equal(out.cpu(), ref)          # gathered values: bit-exact rows in the same (score) order
    assert torch.equal(idx.cpu().long(), ridx)
    refa, ridxa = oops.e2e_postprocess(y, max_det, nc, agnostic=True)
    outa, idxa = E.topk_postprocess(y.cuda(), max_det, nc, agnostic=True)
    assert torch.equal(outa.cpu(), refa) and torch.equal(idxa.cpu().long(), ridxa)


@pytest.mark.gpu
def test_gpu_topk_postprocess_ties_are_a_valid_selection():
    import yolosharp_b200.engine as E
    B, nc, A, k = 2, 6, 500, 50
    y = _pred(B, nc, A, 2, ties=True)
    out, idx = E.topk_postprocess(y.cuda(), k, nc)
    out, idx = out.cpu(), idx.cpu().long()
    ref, _ = oops.e2e_postprocess(y, k, nc)
    # the multiset of selected scores is determined even with ties; rows are score-descending; every row is consistent
    assert torch.equal(out[..., 4].sort(descending=True).values, ref[..., 4].sort(descending=True).values)
    assert bool((out[..., 4][:, 1:] <= out[..., 4][:, :-1]).all())
    for b in range(B):
        for j in range(k):
            a, c = int(idx[b, j]), int(out[b, j, 5])
            assert float(y[b, 4 + c, a]) == float(out[b, j, 4]) and torch.equal(y[b, :4, a], out[b, j, :4])
