"""fp16-emulating oracle (TEST INFRASTRUCTURE ONLY - never imported by the product path).

The reference's fp16 predict path (`Config.ScalarType = Float16`, Models/Detector.cs:41 `.to(dtype, device)`)
runs libtorch's half kernels; the engine's throughput mode (YB_PREC_F16) is *fp16 storage + fp32 accumulate*
with BatchNorm folded into the weights.  Neither is bit-comparable with the fp32 oracle, so this module
restates the fp32 oracle graph (oracle/modules.py, oracle/yolo.py - same wiring, same quirks) with exactly the
rounding points of the engine:

  * BN folded in double, `w*scale` rounded double -> float -> half (csrc/engine.cu finalize_conv), bias fp32;
  * every activation tensor that the engine stores (NHWC fp16) is rounded to fp16 where the engine stores it:
    after `SiLU(conv + bias) [+ residual]` - the Bottleneck / PSABlock shortcut is added in fp32 BEFORE the
    rounding (conv_tc.cu epilogue), the attention output is rounded before `pe(v)` is added to it;
  * fp32 accumulation (torch CPU conv in fp32 on fp16-valued operands: products are exact, sums differ from the
    TMEM accumulator only by summation order ~1e-6);
  * the final 1x1 convs of the Detect branches are NOT rounded (their accumulators go straight through
    DFL / sigmoid into the fp32 prediction tensor); Proto output is fp16-valued.

What it does NOT emulate bit-for-bit: `tanh.approx` inside the engine's SiLU (rel. error <= 2^-11, the size of
one fp16 ulp), `__expf`/`__fdividef` in the decode, and summation order.  Parity tests therefore compare at
~1e-3 of the layer range instead of exactly, and state the fp16 -> fp32-oracle gap separately.
Reference lines restated: Modules/Convs.cs:36-56, Block.cs:572-607 (Bottleneck), :697-722 (PSABlock),
:752-809 (Attention), :51-84 (Proto), Head.cs:71-87.
"""
import copy

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import modules as om


def r16(t):
    """round to fp16 storage, keep computing in fp32"""
    return t.half().float()


def fold_bn(conv, bn):
    """engine.cu finalize_conv: scale = gamma / sqrt(var + 1e-3) in double; w = half(float(w * scale));
    bias = float(beta - mean * scale)."""
    scale = bn.weight.double() / torch.sqrt(bn.running_var.double() + 1e-3)
    w = (conv.weight.double() * scale.view(-1, 1, 1, 1)).float().half().float()
    b = (bn.bias.double() - bn.running_mean.double() * scale).float()
    return w, b


class EConv(nn.Module):
    """Conv + folded BN + act with the engine's rounding.  forward(x, res=None, rnd=True)."""

    def __init__(self, m):
        super().__init__()
        w, b = fold_bn(m.conv, m.bn)
        self.register_buffer("w", w)
        self.register_buffer("b", b)
        self.stride, self.padding, self.groups = m.conv.stride, m.conv.padding, m.conv.groups
        self.silu = isinstance(m.act, nn.SiLU)

    def forward(self, x, res=None, rnd=True):
        z = F.conv2d(x, self.w, None, self.stride, self.padding, 1, self.groups) + self.b.view(1, -1, 1, 1)
        if self.silu:
            z = F.silu(z)
        if res is not None:
            z = z + res
        return r16(z) if rnd else z


class EPlain(nn.Module):
    """nn.Conv2d(bias) of the Detect tails: fp16 weights, fp32 bias, fp32 output (not stored)."""

    def __init__(self, m):
        super().__init__()
        self.register_buffer("w", m.weight.detach().half().float())
        self.register_buffer("b", m.bias.detach().float())

    def forward(self, x):
        return F.conv2d(x, self.w, self.b)


class EConvT(nn.Module):
    """Proto.upsample ConvTranspose2d(c, c, 2, 2, bias): the engine runs it as a 1x1 conv to 4c + pixel shuffle,
    output stored in fp16."""

    def __init__(self, m):
        super().__init__()
        self.register_buffer("w", m.weight.detach().half().float())
        self.register_buffer("b", m.bias.detach().float())

    def forward(self, x):
        return r16(F.conv_transpose2d(x, self.w, self.b, 2, 0))


class EBottleneck(nn.Module):
    def __init__(self, m):
        super().__init__()
        self.cv1, self.cv2, self.add = EConv(m.cv1), EConv(m.cv2), m.add

    def forward(self, x):
        return self.cv2(self.cv1(x), res=x if self.add else None)


class EAttention(nn.Module):
    def __init__(self, m):
        super().__init__()
        self.num_heads, self.head_dim, self.key_dim, self.scale = m.num_heads, m.head_dim, m.key_dim, m.scale
        self.qkv, self.proj, self.pe = EConv(m.qkv), EConv(m.proj), EConv(m.pe)

    def forward(self, x, res=None):
        B, C, H, W = x.shape
        N = H * W
        qkv = self.qkv(x)
        q, k, v = qkv.view(B, self.num_heads, self.key_dim * 2 + self.head_dim, N).split(
            [self.key_dim, self.key_dim, self.head_dim], dim=2)
        attn = (q.transpose(-2, -1).matmul(k) * self.scale).softmax(dim=-1)
        ao = r16(v.matmul(attn.transpose(-2, -1)).view(B, C, H, W))   # attention_kernel stores fp16
        xs = self.pe(v.reshape(B, C, H, W), res=ao)                    # dwconv epilogue adds the stored ao
        return self.proj(xs, res=res)


class EPSABlock(nn.Module):
    def __init__(self, m):
        super().__init__()
        self.attn = EAttention(m.attn)
        self.ffn = nn.ModuleList([EConv(m.ffn[0]), EConv(m.ffn[1])])  # same child names as the oracle's Sequential
        self.add = m.add

    def forward(self, x):
        b1 = self.attn(x, res=x if self.add else None)
        return self.ffn[1](self.ffn[0](b1), res=b1 if self.add else None)


def _convert(mod):
    for name, child in list(mod.named_children()):
        if isinstance(child, om.Bottleneck):
            setattr(mod, name, EBottleneck(child))
        elif isinstance(child, om.PSABlock):
            setattr(mod, name, EPSABlock(child))
        elif isinstance(child, om.Attention):
            setattr(mod, name, EAttention(child))
        elif isinstance(child, om.Conv):
            setattr(mod, name, EConv(child))
        elif isinstance(child, om.DFL):
            continue  # DFL.conv is the fp32 arange(16) expectation (Block.cs:29-30): untouched
        elif isinstance(child, nn.ConvTranspose2d):
            setattr(mod, name, EConvT(child))
        elif isinstance(child, nn.Conv2d):
            setattr(mod, name, EPlain(child))
        else:
            _convert(child)


def convert(model):
    """Deep copy of an oracle model (eval mode) whose forward has the engine's fp16 rounding points.  Module
    names are preserved, so forward hooks give per-layer expectations under the reference names."""
    m = copy.deepcopy(model).eval()
    _convert(m)
    return m


def input_f16(x):
    """network input as the engine's fp16 stem sees a float tensor"""
    return r16(x.float())


def input_u8(u8):
    """uint8 input: stem_tc_kernel computes half(b * half(1/255)) with one rounding (conv_tc.cu stem_load4)"""
    k = torch.tensor(1.0 / 255.0).half().float()
    return r16(u8.float() * k)
