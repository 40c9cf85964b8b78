"""Oracle restatement of the reference graph wiring (test infrastructure only).

Follows Models/Yolo.cs:10-135 (Yolov8), :200-258 (Yolov11), :337-352
(Yolov8Segment) of /root/reference/YoloSharp.  The layer list is held in
``self.model`` so state_dict keys are ``model.{i}....`` like the reference's.
"""
import torch
import torch.nn as nn

from .modules import C2PSA, C2f, C3k2, Conv, Detect, SPPF, Segment

V8_SIZES = {  # Yolo.cs:45-49  (depth_multiple, width_multiple, max_channels)
    "n": (0.34, 0.25, 1024), "s": (0.34, 0.5, 1024), "m": (0.67, 0.75, 576),
    "l": (1.0, 1.0, 512), "x": (1.0, 1.25, 640),
}
V11_SIZES = {  # Yolo.cs:213-217 (+useC3k)
    "n": (0.5, 0.25, 1024, False), "s": (0.5, 0.5, 1024, False), "m": (0.5, 1.0, 512, True),
    "l": (1.0, 1.0, 512, True), "x": (1.0, 1.5, 768, True),
}


class Concat(nn.Module):
    """Modules/Convs.cs:435-448."""

    def forward(self, xs):
        return torch.cat(xs, 1)


class Yolov8(nn.Module):
    output_indexs = (4, 6, 9, 12, 15, 18, 21)  # Yolo.cs:13
    concat_index = (1, 0, 3, 2)  # Yolo.cs:14

    def __init__(self, nc=80, reg_max=16, size="n"):
        super().__init__()
        self.nc, self.reg_max, self.size = nc, reg_max, size
        self.model = nn.ModuleList(self.build_model())

    def widths_depths(self):
        d, w, mc = V8_SIZES[self.size]
        # C# (int)(w * width_multiple) with float32 multipliers; all products are exact here
        widths = [min(int(x * w), mc) for x in (64, 128, 256, 512, 1024)]
        depths = [int(x * d) for x in (3, 6, 9)]
        return widths, depths

    def make_head(self, ch):
        return Detect(self.nc, self.reg_max, ch, legacy=True)

    def build_model(self):
        """Yolo.cs:41-89."""
        w, dp = self.widths_depths()
        self.ch = (w[2], w[3], w[4])
        return [
            Conv(3, w[0], 3, 2),
            Conv(w[0], w[1], 3, 2),
            C2f(w[1], w[1], dp[0], True),
            Conv(w[1], w[2], 3, 2),
            C2f(w[2], w[2], dp[1], True),
            Conv(w[2], w[3], 3, 2),
            C2f(w[3], w[3], dp[1], True),
            Conv(w[3], w[4], 3, 2),
            C2f(w[4], w[4], dp[0], True),
            SPPF(w[4], w[4], 5),
            nn.Upsample(scale_factor=2, mode="nearest"),
            Concat(),
            C2f(w[3] + w[4], w[3], dp[0]),
            nn.Upsample(scale_factor=2, mode="nearest"),
            Concat(),
            C2f(w[2] + w[3], w[2], dp[0]),
            Conv(w[2], w[2], 3, 2),
            Concat(),
            C2f(w[2] + w[3], w[3], dp[0]),
            Conv(w[3], w[3], 3, 2),
            Concat(),
            C2f(w[4] + w[3], w[4], dp[0]),
            self.make_head(self.ch),
        ]

    def forward(self, x):
        """Yolo.cs:92-134."""
        outputs, cat_count, result = [], 0, None
        for i, m in enumerate(self.model):
            if isinstance(m, Concat):
                x = m([x, outputs[self.concat_index[cat_count]]])
                cat_count += 1
            elif isinstance(m, Detect):
                result = m([outputs[-3], outputs[-2], outputs[-1]])
            else:
                x = m(x)
            if i in self.output_indexs:
                outputs.append(x)
        return result


class Yolov11(Yolov8):
    output_indexs = (4, 6, 10, 13, 16, 19, 22)  # Yolo.cs:202

    def make_head(self, ch):
        return Detect(self.nc, self.reg_max, ch, legacy=False)

    def build_model(self):
        """Yolo.cs:209-257."""
        d, wm, mc, c3k = V11_SIZES[self.size]
        w = [min(int(x * wm), mc) for x in (64, 128, 256, 512, 1024)]
        n = int(2 * d)
        self.ch = (w[2], w[3], w[4])
        return [
            Conv(3, w[0], 3, 2),
            Conv(w[0], w[1], 3, 2),
            C3k2(w[1], w[2], n, c3k, e=0.25),
            Conv(w[2], w[2], 3, 2),
            C3k2(w[2], w[3], n, c3k, e=0.25),
            Conv(w[3], w[3], 3, 2),
            C3k2(w[3], w[3], n, c3k=True),
            Conv(w[3], w[4], 3, 2),
            C3k2(w[4], w[4], n, c3k=True),
            SPPF(w[4], w[4], 5),
            C2PSA(w[4], w[4], n),
            nn.Upsample(scale_factor=2, mode="nearest"),
            Concat(),
            C3k2(w[4] + w[3], w[3], n, c3k),
            nn.Upsample(scale_factor=2, mode="nearest"),
            Concat(),
            C3k2(w[3] + w[3], w[2], n, c3k),
            Conv(w[2], w[2], 3, 2),
            Concat(),
            C3k2(w[3] + w[2], w[3], n, c3k),
            Conv(w[3], w[3], 3, 2),
            Concat(),
            C3k2(w[4] + w[3], w[4], n, c3k=True),
            self.make_head(self.ch),
        ]


class Yolov8Segment(Yolov8):
    """Yolo.cs:337-352: Detect replaced by Segment(npr = ch[0])."""

    def make_head(self, ch):
        return Segment(self.nc, 32, ch[0], self.reg_max, ch, legacy=True)


class Yolov11Segment(Yolov11):
    """Yolo.cs:354-370."""

    def make_head(self, ch):
        return Segment(self.nc, 32, ch[0], self.reg_max, ch, legacy=False)


def build(arch="v8", task="detect", size="n", nc=80):
    cls = {("v8", "detect"): Yolov8, ("v11", "detect"): Yolov11,
           ("v8", "segment"): Yolov8Segment, ("v11", "segment"): Yolov11Segment}[(arch, task)]
    return cls(nc=nc, size=size)


def synth_weights(model, seed=0, cls_bias=None, head_gain=1.0):
    """Seeded synthetic weights for sizes with no shipped checkpoint:
    conv ~ N(0, 2/fan_in) scaled down slightly so activations stay O(1) through ~60 layers,
    BN gamma U(0.5,1.5), beta N(0,0.1), running_mean N(0,0.1), running_var U(0.5,1.5);
    values are rounded through fp16 so an fp16-storage engine holds identical weights.
    cls_bias / head_gain: bias and weight gain of the final 1x1 convs (control how many anchors
    pass the conf filter and how peaked the DFL distributions are)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, mod in model.named_modules():
            if isinstance(mod, (nn.Conv2d, nn.ConvTranspose2d)):
                if name.endswith("dfl.conv"):
                    continue
                w = mod.weight
                fan_in = w.shape[1] * w.shape[2] * w.shape[3]
                w.copy_((torch.randn(w.shape, generator=g) * (1.6 / fan_in) ** 0.5).half().float())
                if mod.bias is not None:
                    mod.bias.copy_((torch.randn(mod.bias.shape, generator=g) * 0.1).half().float())
            elif isinstance(mod, nn.BatchNorm2d):
                c = mod.num_features
                mod.weight.copy_((torch.rand(c, generator=g) + 0.5).half().float())
                mod.bias.copy_((torch.randn(c, generator=g) * 0.1).half().float())
                mod.running_mean.copy_((torch.randn(c, generator=g) * 0.1).half().float())
                mod.running_var.copy_((torch.rand(c, generator=g) + 0.5).half().float())
        head = model.model[-1]
        if head_gain != 1.0:  # spread the final logits so that NMS sees a realistic candidate set
            for branch in (head.cv2, head.cv3):
                for seq in branch:
                    seq[-1].weight.copy_((seq[-1].weight * head_gain).half().float())
        if cls_bias is not None:
            for seq in head.cv3:
                seq[-1].bias.fill_(float(torch.tensor(cls_bias).half()))
    return model
