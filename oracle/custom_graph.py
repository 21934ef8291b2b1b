"""Oracle: fp32 PyTorch restatement of the reference's hand-written network
``unina_yolo_dla/model.py`` (``UNINA_YOLO_DLA``).  TEST INFRASTRUCTURE.

PINNED: ``tests/test_oracle_pins.py`` checks this restatement bit-for-bit against the
real module imported from /root/reference (when mounted) and against
tests/golden/custom_*.npz, which tests/golden/make_golden.py produced by running the
real module in this container.

Same ``state_dict`` keys as the reference (378 entries):
``backbone.{stem,stage1_conv,stage1_block,stage2_conv,stage2_c3k2,stage3_conv,
stage3_c3k2,sppf}.*``, ``neck.{lateral_p3,fpn_c3k2_1,lateral_p2,fpn_c3k2_2,down1,
pan_c3k2_1,down2,pan_c3k2_2}.*``, ``head_p{2,3,4}.{cls,reg}_branch.{0,1}.{conv,bn}.*``
and ``....2.{weight,bias}``.

Arithmetic followed (file:line in /root/reference/unina_yolo_dla/model.py):
ConvBlock 23-50 (bias-free conv, BN eps 1e-5, ReLU), Bottleneck 53-73 (1x1 -> 3x3,
residual when widths match), C3k2 76-110 (two 1x1 paths, n bottlenecks, cat, 1x1),
SPPF_DLA 113-132, nearest Upsample 135-147, Backbone 152-219, Neck 224-269,
DetectionHead 274-303, UNINA_YOLO_DLA.forward 347-365.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class CBR(nn.Module):
    def __init__(self, cin, cout, k=3, s=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, s, k // 2, bias=False)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class Res(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.cv1 = CBR(c, c, 1)
        self.cv2 = CBR(c, c, 3)

    def forward(self, x):
        return x + self.cv2(self.cv1(x))


class CSP(nn.Module):
    def __init__(self, cin, cout, n=1):
        super().__init__()
        h = cout // 2
        self.cv1 = CBR(cin, h, 1)
        self.cv2 = CBR(cin, h, 1)
        self.bottlenecks = nn.Sequential(*(Res(h) for _ in range(n)))
        self.cv3 = CBR(2 * h, cout, 1)

    def forward(self, x):
        return self.cv3(torch.cat((self.bottlenecks(self.cv1(x)), self.cv2(x)), 1))


class PoolPyramid(nn.Module):
    def __init__(self, cin, cout, k=5):
        super().__init__()
        self.cv1 = CBR(cin, cin // 2, 1)
        self.cv2 = CBR(cin * 2, cout, 1)
        self.k = k

    def forward(self, x):
        outs = [self.cv1(x)]
        for _ in range(3):
            outs.append(F.max_pool2d(outs[-1], self.k, 1, self.k // 2))
        return self.cv2(torch.cat(outs, 1))


class _Backbone(nn.Module):
    def __init__(self, bc, lite_p2):
        super().__init__()
        c = [bc * m for m in (1, 2, 4, 8)]
        self.stem = CBR(3, c[0], 3, 2)
        self.stage1_conv = CBR(c[0], c[1], 3, 2)
        self.stage1_block = CBR(c[1], c[1], 3) if lite_p2 else CSP(c[1], c[1], 1)
        self.stage2_conv = CBR(c[1], c[2], 3, 2)
        self.stage2_c3k2 = CSP(c[2], c[2], 2)
        self.stage3_conv = CBR(c[2], c[3], 3, 2)
        self.stage3_c3k2 = CSP(c[3], c[3], 2)
        self.sppf = PoolPyramid(c[3], c[3])
        self.widths = c[1:]

    def forward(self, x):
        p2 = self.stage1_block(self.stage1_conv(self.stem(x)))
        p3 = self.stage2_c3k2(self.stage2_conv(p2))
        p4 = self.stage3_c3k2(self.stage3_conv(p3))
        return p2, p3, p4, self.sppf(p4)


class _Neck(nn.Module):
    def __init__(self, c2, c3, c4):
        super().__init__()
        self.lateral_p3 = CBR(c4, c3, 1)
        self.fpn_c3k2_1 = CSP(2 * c3, c3, 1)
        self.lateral_p2 = CBR(c3, c2, 1)
        self.fpn_c3k2_2 = CSP(2 * c2, c2, 1)
        self.down1 = CBR(c2, c2, 3, 2)
        self.pan_c3k2_1 = CSP(c2 + c3, c3, 1)
        self.down2 = CBR(c3, c3, 3, 2)
        self.pan_c3k2_2 = CSP(c3 + c4, c4, 1)

    def forward(self, feats):
        p2, p3, p4, ctx = feats
        up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")
        f3 = self.fpn_c3k2_1(torch.cat((up(self.lateral_p3(ctx)), p3), 1))
        f2 = self.fpn_c3k2_2(torch.cat((up(self.lateral_p2(f3)), p2), 1))
        o3 = self.pan_c3k2_1(torch.cat((self.down1(f2), f3), 1))
        o4 = self.pan_c3k2_2(torch.cat((self.down2(o3), p4), 1))
        return f2, o3, o4


class _Head(nn.Module):
    def __init__(self, c, nc):
        super().__init__()
        self.cls_branch = nn.Sequential(CBR(c, c, 3), CBR(c, c, 3), nn.Conv2d(c, nc, 1))
        self.reg_branch = nn.Sequential(CBR(c, c, 3), CBR(c, c, 3), nn.Conv2d(c, 4, 1))

    def forward(self, x):
        return self.cls_branch(x), self.reg_branch(x)


class CustomNet(nn.Module):
    """``UNINA_YOLO_DLA(num_classes, base_channels, lite_p2)`` restated."""

    def __init__(self, num_classes=4, base_channels=32, lite_p2=False):
        super().__init__()
        self.num_classes = num_classes
        self.backbone = _Backbone(base_channels, lite_p2)
        self.neck = _Neck(*self.backbone.widths)
        c2, c3, c4 = self.backbone.widths
        self.head_p2 = _Head(c2, num_classes)
        self.head_p3 = _Head(c3, num_classes)
        self.head_p4 = _Head(c4, num_classes)

    def forward(self, x):
        f2, f3, f4 = self.neck(self.backbone(x))
        return [self.head_p2(f2), self.head_p3(f3), self.head_p4(f4)]
