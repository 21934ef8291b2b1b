"""Oracle: fp32 PyTorch restatement of the Ultralytics graph the reference builds from
``unina_yolo_dla/unina-yolo-dla-m.yaml``.  TEST INFRASTRUCTURE -- never imported by the
product package.

PARITY UNPINNED: ``ultralytics`` (requirements.txt:1, unpinned; >= 8.3.0 needed for
C3k2) is not vendored under /root/reference and cannot be installed here.  Everything
below restates the published behaviour of Ultralytics 8.3.x and is anchored on the
reference's own call sites:

* construction  : trainer.py:146-158 (``DetectionModel(cfg, nc)``), trainer.py:85-94
                  (scale/scales injection), unina-yolo-dla-m.yaml:14-62
* SPPF_DLA      : trainer.py:108-124 (incl. the ``c2 == k and c2 < 16`` arg repair)
* SiLU -> ReLU  : trainer.py:130-136
* Detect attrs  : mine_data.py:110-161 (``nl``, ``cv2[i]`` = box, ``cv3[i]`` = cls)
* output        : train.py:396-423, trainer.py:237-240 ([N,6] = x1,y1,x2,y2,conf,cls)

Assumed third-party behaviours (re-check against a real install when one exists):
``round(n*depth)`` banker's rounding; ``scale in "mlx"`` forces ``c3k=True`` and marks the
model non-legacy (depth-wise cls branch in Detect); BN ``eps=1e-3, momentum=0.03``;
``Detect.bias_init`` (box bias 1.0, cls bias ``log(5/nc/(640/s)^2)``); strides [4,8,16]
from a dry run; ``make_anchors`` offset 0.5, row-major, levels P2->P3->P4.

The module tree reproduces the reference ``state_dict`` key schema exactly
(925 entries).
"""
from __future__ import annotations

import ast
import math
from math import gcd
from pathlib import Path

import torch
import torch.nn as nn
import yaml

BN_EPS = 1e-3
BN_MOMENTUM = 0.03


def make_divisible(x, divisor):
    return int(math.ceil(x / divisor) * divisor)


class Conv(nn.Module):
    """Conv2d(bias=False, pad=k//2) -> BatchNorm2d -> ReLU  (act swapped by trainer.py:130-136)."""

    def __init__(self, c1, c2, k=1, s=1, g=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, groups=g, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=BN_EPS, momentum=BN_MOMENTUM)
        self.act = nn.ReLU(inplace=True) if act else nn.Identity()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class DWConv(Conv):
    def __init__(self, c1, c2, k=1, s=1):
        super().__init__(c1, c2, k, s, g=gcd(c1, c2))


class Bottleneck(nn.Module):
    def __init__(self, c1, c2, shortcut=True, g=1, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0], 1)
        self.cv2 = Conv(c_, c2, k[1], 1, g=g)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C3k(nn.Module):
    """C3 with k x k bottlenecks: cv3(cat(m(cv1(x)), cv2(x)))."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5, k=3):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, g, k=(k, k), e=1.0) for _ in range(n)))

    def forward(self, x):
        return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), 1))


class C3k2(nn.Module):
    """C2f whose inner blocks are C3k (c3k=True) or Bottleneck."""

    def __init__(self, c1, c2, n=1, c3k=False, e=0.5, g=1, shortcut=True):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(
            C3k(self.c, self.c, 2, shortcut, g) if c3k else Bottleneck(self.c, self.c, shortcut, g)
            for _ in range(n)
        )

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        y.extend(m(y[-1]) for m in self.m)
        return self.cv2(torch.cat(y, 1))


class SPPF_DLA(nn.Module):
    """trainer.py:108-124: 1x1 -> three cascaded MaxPool(k,1,k//2) -> cat(4) -> 1x1."""

    def __init__(self, c1, c2, k=5):
        super().__init__()
        if c2 == k and c2 < 16:  # trainer.py:112-113 (parse_model passes (c1, 5))
            c2 = c1
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        x = self.cv1(x)
        y1 = self.m(x)
        y2 = self.m(y1)
        y3 = self.m(y2)
        return self.cv2(torch.cat((x, y1, y2, y3), 1))


class Concat(nn.Module):
    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, xs):
        return torch.cat(xs, self.d)


class DFL(nn.Module):
    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1

    def forward(self, x):
        b, _, a = x.shape
        return self.conv(x.view(b, 4, self.c1, a).transpose(2, 1).softmax(1)).view(b, 4, a)


def make_anchors(feats, strides, offset=0.5):
    pts, st = [], []
    for f, s in zip(feats, strides):
        h, w = f.shape[2:]
        sx = torch.arange(w, dtype=torch.float32) + offset
        sy = torch.arange(h, dtype=torch.float32) + offset
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s)))
    return torch.cat(pts), torch.cat(st)


def dist2bbox(distance, anchor_points, xywh=True, dim=-1):
    lt, rb = distance.chunk(2, dim)
    x1y1 = anchor_points - lt
    x2y2 = anchor_points + rb
    if xywh:
        return torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), dim)
    return torch.cat((x1y1, x2y2), dim)


class Detect(nn.Module):
    """Anchor-free DFL head, non-legacy (depth-wise cls branch)."""

    def __init__(self, nc=80, ch=()):
        super().__init__()
        self.nc = nc
        self.nl = len(ch)
        self.reg_max = 16
        self.no = nc + self.reg_max * 4
        self.stride = torch.zeros(self.nl)
        c2 = max(16, ch[0] // 4, self.reg_max * 4)
        c3 = max(ch[0], min(nc, 100))
        self.cv2 = nn.ModuleList(
            nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch
        )
        self.cv3 = nn.ModuleList(
            nn.Sequential(
                nn.Sequential(DWConv(x, x, 3), Conv(x, c3, 1)),
                nn.Sequential(DWConv(c3, c3, 3), Conv(c3, c3, 1)),
                nn.Conv2d(c3, nc, 1),
            )
            for x in ch
        )
        self.dfl = DFL(self.reg_max)

    def raw(self, feats):
        return [torch.cat((self.cv2[i](f), self.cv3[i](f)), 1) for i, f in enumerate(feats)]

    def decode(self, x):
        b = x[0].shape[0]
        x_cat = torch.cat([xi.reshape(b, self.no, -1) for xi in x], 2)
        anchors, strides = (t.transpose(0, 1) for t in make_anchors(x, self.stride, 0.5))
        box, cls = x_cat.split((self.reg_max * 4, self.nc), 1)
        dbox = dist2bbox(self.dfl(box), anchors.unsqueeze(0), xywh=True, dim=1) * strides
        return torch.cat((dbox, cls.sigmoid()), 1)

    def forward(self, feats):
        x = self.raw(list(feats))
        if self.training:
            return x
        return self.decode(x), x

    def bias_init(self):
        for a, b, s in zip(self.cv2, self.cv3, self.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[: self.nc] = math.log(5 / self.nc / (640 / float(s)) ** 2)


def _literal(a):
    """Ultralytics evaluates string args (``"None"`` -> None; ``"nearest"`` stays a string)."""
    if isinstance(a, str):
        try:
            return ast.literal_eval(a)
        except (ValueError, SyntaxError):
            return a
    return a


def parse_model(d, ch=3):
    """Restated ``parse_model`` for the module set the YAML uses."""
    nc = d["nc"]
    scale = d.get("scale", "m")  # trainer.py:89-90
    scales = d.get("scales", {scale: [1.0, 1.0, 1024]})  # trainer.py:91-93
    depth, width, max_ch = scales[scale]
    chs = [ch]
    layers, save = [], []
    for i, (f, n, m, args) in enumerate(d["backbone"] + d["head"]):
        args = [nc if a == "nc" else _literal(a) for a in args]
        n = max(round(n * depth), 1) if n > 1 else n
        if m in ("Conv", "C3k2"):
            c1, c2 = chs[f], args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_ch) * width, 8)
            args = [c1, c2, *args[1:]]
            if m == "C3k2":
                args.insert(2, n)
                n = 1
                if scale in "mlx":
                    args[3] = True
            mod = Conv(*args) if m == "Conv" else C3k2(*args)
        elif m == "SPPF_DLA":  # not a "base module": args pass through unchanged
            c2 = chs[f]
            mod = SPPF_DLA(*args_for_sppf(chs[f], args))
        elif m == "nn.Upsample":
            c2 = chs[f]
            mod = nn.Upsample(*args)
        elif m == "Concat":
            c2 = sum(chs[x] for x in f)
            mod = Concat(*args)
        elif m == "Detect":
            args.append([chs[x] for x in f])
            c2 = None
            mod = Detect(*args)
        else:
            raise ValueError(f"module {m} is not part of the restated set")
        mod.i, mod.f = i, f
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
        layers.append(mod)
        if i == 0:
            chs = []
        chs.append(c2)
    return nn.Sequential(*layers), sorted(save)


def args_for_sppf(c_in, args):
    """Ultralytics does not know SPPF_DLA, so it calls ``SPPF_DLA(*args)`` with the raw YAML
    args ``[128, 5]`` i.e. (c1=128, c2=5); trainer.py:112-113 then repairs c2 := c1 and the
    default k=5 applies.  ``c_in`` is only used to assert the YAML is self-consistent."""
    assert args[0] == c_in, "YAML SPPF_DLA arg 0 must equal its input width"
    return args


class DetectionModel(nn.Module):
    """``DetectionModel(cfg, nc)`` + ``replace_silu_with_relu`` (trainer.py:156-158)."""

    def __init__(self, cfg, ch=3, nc=None):
        super().__init__()
        self.yaml = cfg if isinstance(cfg, dict) else yaml.safe_load(Path(cfg).read_text())
        self.yaml = dict(self.yaml)
        if nc is not None:
            self.yaml["nc"] = nc
        self.model, self.save = parse_model(self.yaml, ch)
        self.nc = self.yaml["nc"]
        self.names = {i: str(i) for i in range(self.nc)}
        det = self.model[-1]
        det.stride = torch.tensor([4.0, 8.0, 16.0])  # dry-run result for this graph
        self.stride = det.stride
        det.bias_init()

    def forward_features(self, x, want=None):
        """Runs the routed layer loop; returns (last output, {layer index: output})."""
        y, kept = [], {}
        for m in self.model:
            if m.f != -1:
                x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
            x = m(x)
            y.append(x if m.i in self.save else None)
            if want is not None and m.i in want:
                kept[m.i] = x
        return x, kept

    def forward(self, x):
        return self.forward_features(x)[0]


def default_yaml_path():
    """The YAML ships with the product package (a config, not code); the reference copy is
    used when the mount exists so that drift is caught."""
    here = Path(__file__).resolve().parent.parent / "unina-yolo-dla_b200" / "unina-yolo-dla-m.yaml"
    return here
