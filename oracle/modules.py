"""Oracle restatement of the reference NN modules (test infrastructure only).

Each class cites the reference C# it follows (paths relative to
/root/reference/YoloSharp).  Attribute names equal the reference's so that
``state_dict()`` keys equal the reference's ``.bin`` tensor names.
Quirks of the reference are reproduced on purpose (see SURVEY.md §8 notes).
"""
import math

import torch
import torch.nn as nn


class Conv(nn.Module):
    """Conv2d(bias=False) + BatchNorm2d(eps 1e-3, momentum 0.03) + SiLU.
    Modules/Convs.cs:36-56 (pad = k/2, never fused: forward_fuse is dead code)."""

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=None):
        super().__init__()
        p = k // 2 if p is None else p
        self.conv = nn.Conv2d(c1, c2, k, s, p, groups=g, bias=False, dilation=d)
        self.bn = nn.BatchNorm2d(c2, eps=0.001, momentum=0.03)
        self.act = act if act is not None else nn.SiLU()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class DWConv(Conv):
    """Modules/Convs.cs:108-114: groups = gcd(c1, c2)."""

    def __init__(self, c1, c2, k=1, s=1, d=1, act=None):
        super().__init__(c1, c2, k, s, g=math.gcd(c1, c2), d=d, act=act)


class Bottleneck(nn.Module):
    """Modules/Block.cs:572-607."""

    def __init__(self, c1, c2, shortcut=True, g=1, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0], 1)
        self.cv2 = Conv(c_, c2, k[1], 1, g=g)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        return x + self.cv2(self.cv1(x)) if self.add else self.cv2(self.cv1(x))


class C2f(nn.Module):
    """Modules/Block.cs:371-398 (Bottlenecks built with e=1.0)."""

    def __init__(self, c1, c2, n=1, shortcut=False, g=1, e=0.5):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, self.c, shortcut, g, k=(3, 3), e=1.0) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        for m in self.m:
            y.append(m(y[-1]))
        return self.cv2(torch.cat(y, 1))


class C3(nn.Module):
    """Modules/Block.cs:404-441."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, g, k=(1, 3), e=1.0) for _ in range(n)))

    def forward(self, x):
        return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), 1))


class C3k(C3):
    """Modules/Block.cs:611-620: C3 whose bottlenecks are k=(3,3), e=1.0."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__(c1, c2, n, shortcut, g, e)
        c = int(c2 * e)
        self.m = nn.Sequential(*(Bottleneck(c, c, shortcut, g, k=(3, 3), e=1.0) for _ in range(n)))


class C3k2(nn.Module):
    """Modules/Block.cs:623-661.  Non-c3k members are Bottleneck with the default e=0.5."""

    def __init__(self, c1, c2, n=1, c3k=False, e=0.5, g=1, shortcut=True):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(
            C3k(self.c, self.c, 2, shortcut, g) if c3k else Bottleneck(self.c, self.c, shortcut, g, k=(3, 3))
            for _ in range(n)
        )

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        for m in self.m:
            y.append(m(y[-1]))
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    """Modules/Block.cs:236-282.  cv1 has act=Identity (reference quirk, :257)."""

    def __init__(self, c1, c2, k=5, n=3, shortcut=False):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1, act=nn.Identity())
        self.cv2 = Conv(c_ * (n + 1), c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)
        self.n = n
        self.add = shortcut and c1 == c2

    def forward(self, x):
        y = [self.cv1(x)]
        for _ in range(self.n):
            y.append(self.m(y[-1]))
        r = self.cv2(torch.cat(y, 1))
        return r + x if self.add else r


class Attention(nn.Module):
    """Modules/Block.cs:752-809 (SelfAttention branch).  qkv/proj/pe keep the
    default SiLU (reference quirk; Ultralytics uses act=False)."""

    def __init__(self, dim, num_heads=8, attn_ratio=0.5):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.key_dim = int(self.head_dim * attn_ratio)
        self.scale = float(self.key_dim ** -0.5)
        nh_kd = self.key_dim * num_heads
        h = dim + nh_kd * 2
        self.qkv = Conv(dim, h, 1)
        self.proj = Conv(dim, dim, 1)
        self.pe = Conv(dim, dim, 3, 1, g=dim)

    def forward(self, x):
        B, C, H, W = x.shape
        N = H * W
        qkv = self.qkv(x)
        q, k, v = qkv.view(B, self.num_heads, self.key_dim * 2 + self.head_dim, N).split(
            [self.key_dim, self.key_dim, self.head_dim], dim=2)
        attn = q.transpose(-2, -1).matmul(k) * self.scale
        attn = attn.softmax(dim=-1)
        x = v.matmul(attn.transpose(-2, -1)).view(B, C, H, W) + self.pe(v.reshape(B, C, H, W))
        return self.proj(x)


class PSABlock(nn.Module):
    """Modules/Block.cs:697-722; ffn[1] keeps SiLU (reference quirk)."""

    def __init__(self, c, attn_ratio=0.5, num_heads=8, shortcut=True):
        super().__init__()
        self.attn = Attention(c, num_heads, attn_ratio)
        self.ffn = nn.Sequential(Conv(c, c * 2, 1), Conv(c * 2, c, 1))
        self.add = shortcut

    def forward(self, x):
        x = x + self.attn(x) if self.add else self.attn(x)
        x = x + self.ffn(x) if self.add else self.ffn(x)
        return x


class C2PSA(nn.Module):
    """Modules/Block.cs:664-695."""

    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        assert c1 == c2
        self.c = int(c1 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv(2 * self.c, c2, 1)
        self.m = nn.Sequential(*(PSABlock(self.c, attn_ratio=0.5, num_heads=self.c // 64) for _ in range(n)))

    def forward(self, x):
        a, b = self.cv1(x).split([self.c, self.c], dim=1)
        b = self.m(b)
        return self.cv2(torch.cat((a, b), 1))


class DFL(nn.Module):
    """Modules/Block.cs:15-45: softmax over the 16 bins, expectation via a 1x1
    conv whose weight is arange(16) (always f32)."""

    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float32).view(1, c1, 1, 1)
        self.c1 = c1

    def forward(self, x):
        b, _, a = x.shape
        return self.conv(x.view(b, 4, self.c1, a).transpose(2, 1).softmax(1)).view(b, 4, a)


class Proto(nn.Module):
    """Modules/Block.cs:51-84."""

    def __init__(self, c1, c_=256, c2=32):
        super().__init__()
        self.cv1 = Conv(c1, c_, k=3)
        self.upsample = nn.ConvTranspose2d(c_, c_, 2, 2, 0, bias=True)
        self.cv2 = Conv(c_, c_, k=3)
        self.cv3 = Conv(c_, c2, k=1)

    def forward(self, x):
        return self.cv3(self.cv2(self.upsample(self.cv1(x))))


def make_anchors(feats, strides, grid_cell_offset=0.5):
    """Utils/Tal.cs:313-335."""
    anchor_points, stride_tensor = [], []
    dtype, device = feats[0].dtype, feats[0].device
    for i, stride in enumerate(strides):
        h, w = feats[i].shape[2], feats[i].shape[3]
        sx = torch.arange(w, device=device, dtype=dtype) + grid_cell_offset
        sy = torch.arange(h, device=device, dtype=dtype) + grid_cell_offset
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        anchor_points.append(torch.stack((sx, sy), -1).view(-1, 2))
        stride_tensor.append(torch.full((h * w, 1), stride, dtype=dtype, device=device))
    return torch.cat(anchor_points), torch.cat(stride_tensor)


def dist2bbox(distance, anchor_points, xywh=True, dim=-1):
    """Utils/Tal.cs:338-356."""
    lt, rb = distance.chunk(2, dim)
    x1y1 = anchor_points - lt
    x2y2 = anchor_points + rb
    if xywh:
        c_xy = (x1y1 + x2y2) / 2
        wh = x2y2 - x1y1
        return torch.cat((c_xy, wh), dim)
    return torch.cat((x1y1, x2y2), dim)


class Detect(nn.Module):
    """Modules/Head.cs:8-236 (end2end=False path).  legacy=True -> v8 cls branch
    (two 3x3 Convs); legacy=False -> v11 branch (DW3x3 + 1x1 pairs)."""

    def __init__(self, nc=80, reg_max=16, ch=(), legacy=True):
        super().__init__()
        self.nc = nc
        self.nl = len(ch)
        self.reg_max = reg_max
        self.no = nc + reg_max * 4
        self.stride = [8, 16, 32]
        c2 = max(16, ch[0] // 4, reg_max * 4)
        c3 = max(ch[0], min(nc, 100))
        self.cv2 = nn.ModuleList(
            nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * reg_max, 1)) for x in ch)
        if legacy:
            self.cv3 = nn.ModuleList(
                nn.Sequential(Conv(x, c3, 3), Conv(c3, c3, 3), nn.Conv2d(c3, nc, 1)) for x in ch)
        else:
            self.cv3 = nn.ModuleList(
                nn.Sequential(
                    nn.Sequential(DWConv(x, x, 3), Conv(x, c3, 1)),
                    nn.Sequential(DWConv(c3, c3, 3), Conv(c3, c3, 1)),
                    nn.Conv2d(c3, nc, 1),
                ) for x in ch)
        self.dfl = DFL(reg_max)
        # Head.cs:11-12: tensor fields -> TorchSharp registers them as (empty) buffers,
        # so the shipped .bin files carry `model.N.anchors` / `model.N.strides` of shape [0].
        self.register_buffer("anchors", torch.empty(0))
        self.register_buffer("strides", torch.empty(0))

    def forward_head(self, x):
        """Head.cs:71-87."""
        bs = x[0].shape[0]
        boxes = torch.cat([self.cv2[i](x[i]).view(bs, 4 * self.reg_max, -1) for i in range(self.nl)], dim=-1)
        scores = torch.cat([self.cv3[i](x[i]).view(bs, self.nc, -1) for i in range(self.nl)], dim=-1)
        return {"feats": x, "boxes": boxes, "scores": scores}

    def _get_decode_boxes(self, p):
        """Head.cs:210-223."""
        anchors, strides = make_anchors(p["feats"], self.stride, 0.5)
        anchors, strides = anchors.transpose(0, 1), strides.transpose(0, 1)
        return dist2bbox(self.dfl(p["boxes"]), anchors.unsqueeze(0), xywh=True, dim=1) * strides

    def _inference(self, p):
        """Head.cs:204-208."""
        return torch.cat((self._get_decode_boxes(p), p["scores"].sigmoid()), 1)

    def forward(self, x):
        preds = self.forward_head(list(x))
        if self.training:
            return None, preds
        return {"boxes": self._inference(preds)}, preds


class Segment(Detect):
    """Modules/Head.cs:238-374 (end2end=False path)."""

    def __init__(self, nc=80, nm=32, npr=256, reg_max=16, ch=(), legacy=True):
        super().__init__(nc, reg_max, ch, legacy)
        self.nm, self.npr = nm, npr
        self.proto = Proto(ch[0], npr, nm)
        c4 = max(ch[0] // 4, nm)
        self.cv4 = nn.ModuleList(
            nn.Sequential(Conv(x, c4, 3), Conv(c4, c4, 3), nn.Conv2d(c4, nm, 1)) for x in ch)

    def forward_head(self, x):
        preds = super().forward_head(x)
        bs = x[0].shape[0]
        preds["mask_coefficient"] = torch.cat(
            [self.cv4[i](x[i]).view(bs, self.nm, -1) for i in range(self.nl)], 2)
        return preds

    def _inference(self, p):
        return torch.cat((super()._inference(p), p["mask_coefficient"]), dim=1)

    def forward(self, x):
        inference, preds = super().forward(x)
        proto = self.proto(x[0])
        preds["proto"] = proto
        if self.training:
            return None, preds
        inference["proto"] = proto
        return inference, preds
