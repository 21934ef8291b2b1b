"""``UninaCustomB200``: the reference's hand-written network ``UNINA_YOLO_DLA`` (model.py:308-365)
as a drop-in module on the libuyd plan: same constructor arguments, same 378-entry
``state_dict`` (``backbone.*``, ``neck.*``, ``head_p{2,3,4}.{cls,reg}_branch.*``), same forward
output ``[(cls[B,nc,H,W], reg[B,4,H,W]) x 3]``; ``predict`` adds the TLBR decode and the greedy
class-aware NMS of the reference's post-processing (postprocess.hpp:44-145,
gpu_postprocess.h:42-80) and returns per-image ``[N,6]`` rows (x1,y1,x2,y2,conf,cls).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import UYD_F32, Detection, check
from .plan import NETWORK_INPUT, Plan, Slice, fold_bn


class ConvBlock(nn.Module):
    """model.py:23-50 -- parameters under ``conv.*`` / ``bn.*`` (BN eps 1e-5)."""

    def __init__(self, cin, cout, k=3, s=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, s, k // 2, bias=False)
        self.bn = nn.BatchNorm2d(cout)
        self.act = nn.ReLU(inplace=True)
        self.k, self.s, self.cout = k, s, cout

    def emit(self, p: Plan, src: Slice, dst: Slice | None = None, res: Slice | None = None) -> Slice:
        if src.buf < 0:
            oh, ow = p.in_hw[0] // self.s, p.in_hw[1] // self.s
        else:
            oh, ow = (src.h - 1) // self.s + 1, (src.w - 1) // self.s + 1
        dst = dst or p.buffer(oh, ow, self.cout)
        w, b = fold_bn(self.conv, self.bn)
        return p.conv(src, dst, w, b, self.k, self.s, relu=True, res=res)


class Bottleneck(nn.Module):
    """model.py:53-73 with expansion 1.0: 1x1 -> 3x3, residual."""

    def __init__(self, c):
        super().__init__()
        self.cv1 = ConvBlock(c, c, 1)
        self.cv2 = ConvBlock(c, c, 3)
        self.add = True

    def emit(self, p, src, dst=None):
        return self.cv2.emit(p, self.cv1.emit(p, src), dst, res=src)


class C3k2(nn.Module):
    """model.py:76-110: cv3(cat(bottlenecks(cv1(x)), cv2(x)))."""

    def __init__(self, cin, cout, n=1):
        super().__init__()
        h = int(cout * 0.5)
        self.h = h
        self.cv1 = ConvBlock(cin, h, 1)
        self.cv2 = ConvBlock(cin, h, 1)
        self.bottlenecks = nn.Sequential(*(Bottleneck(h) for _ in range(n)))
        self.cv3 = ConvBlock(2 * h, cout, 1)

    def emit(self, p, src, dst=None):
        cat = p.buffer(src.h, src.w, 2 * self.h)
        t = self.cv1.emit(p, src)
        for i, b in enumerate(self.bottlenecks):
            t = b.emit(p, t, cat.sub(0, self.h) if i == len(self.bottlenecks) - 1 else None)
        self.cv2.emit(p, src, cat.sub(self.h, self.h))
        return self.cv3.emit(p, cat, dst)


class SPPF_DLA(nn.Module):
    """model.py:113-132."""

    def __init__(self, cin, cout, k=5):
        super().__init__()
        self.h = cin // 2
        self.cv1 = ConvBlock(cin, self.h, 1)
        self.cv2 = ConvBlock(self.h * 4, cout, 1)
        self.pool = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def emit(self, p, src, dst=None):
        cat = p.buffer(src.h, src.w, 4 * self.h)
        self.cv1.emit(p, src, cat.sub(0, self.h))
        p.sppf_pool(cat, self.h)
        return self.cv2.emit(p, cat, dst)


class Backbone(nn.Module):
    def __init__(self, base_channels=32, lite_p2=False):
        super().__init__()
        c1, c2, c3, c4 = (base_channels * m for m in (1, 2, 4, 8))
        self.lite_p2 = lite_p2
        self.stem = ConvBlock(3, c1, 3, 2)
        self.stage1_conv = ConvBlock(c1, c2, 3, 2)
        self.stage1_block = ConvBlock(c2, c2, 3) if lite_p2 else C3k2(c2, c2, 1)
        self.stage2_conv = ConvBlock(c2, c3, 3, 2)
        self.stage2_c3k2 = C3k2(c3, c3, 2)
        self.stage3_conv = ConvBlock(c3, c4, 3, 2)
        self.stage3_c3k2 = C3k2(c4, c4, 2)
        self.sppf = SPPF_DLA(c4, c4)
        self.out_channels = [c2, c3, c4]


class Neck(nn.Module):
    def __init__(self, chans):
        super().__init__()
        c2, c3, c4 = chans
        self.lateral_p3 = ConvBlock(c4, c3, 1)
        self.fpn_c3k2_1 = C3k2(c3 * 2, c3, 1)
        self.lateral_p2 = ConvBlock(c3, c2, 1)
        self.fpn_c3k2_2 = C3k2(c2 * 2, c2, 1)
        self.down1 = ConvBlock(c2, c2, 3, 2)
        self.pan_c3k2_1 = C3k2(c2 + c3, c3, 1)
        self.down2 = ConvBlock(c3, c3, 3, 2)
        self.pan_c3k2_2 = C3k2(c3 + c4, c4, 1)
        self.out_channels = [c2, c3, c4]


class DetectionHead(nn.Module):
    """model.py:274-303: decoupled cls (nc) / reg (4, TLBR) branches."""

    def __init__(self, c, nc):
        super().__init__()
        self.cls_branch = nn.Sequential(ConvBlock(c, c, 3), ConvBlock(c, c, 3), nn.Conv2d(c, nc, 1))
        self.reg_branch = nn.Sequential(ConvBlock(c, c, 3), ConvBlock(c, c, 3), nn.Conv2d(c, 4, 1))

    def emit(self, p, f, head: Slice, nc: int):
        for branch, dst in ((self.cls_branch, head.sub(0, nc)), (self.reg_branch, head.sub(nc, 4))):
            t = branch[1].emit(p, branch[0].emit(p, f))
            p.conv(t, dst, branch[2].weight.detach().float().cpu().numpy(), branch[2].bias.detach().float().cpu().numpy(),
                   1, 1, relu=False)


class UninaCustomB200(nn.Module):
    STRIDES = (4, 8, 16)

    def __init__(self, num_classes: int = 4, base_channels: int = 32, lite_p2: bool = False):
        super().__init__()
        self.num_classes = num_classes
        self.backbone = Backbone(base_channels, lite_p2)
        self.neck = Neck(self.backbone.out_channels)
        for lvl, c in zip((2, 3, 4), self.neck.out_channels):
            setattr(self, f"head_p{lvl}", DetectionHead(c, num_classes))
        self._plans = {}
        self._post = {}
        self.eval()

    def refresh(self):
        self._plans.clear()

    @torch.no_grad()
    def init_synthetic(self, seed: int = 0, reg_bias: float = 2.0, gain: float = 1.0) -> "UninaCustomB200":
        """Seeded data-free init (see synth.py); a positive reg bias keeps the TLBR boxes from
        inverting (the raw reg output is signed, model.py:296-300)."""
        from .synth import synthetic_init_

        synthetic_init_(self, seed, gain)
        for lvl in (2, 3, 4):
            getattr(self, f"head_p{lvl}").reg_branch[2].bias.fill_(reg_bias)
        self.refresh()
        return self

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        out = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self.refresh()
        return out

    def _apply(self, fn, recurse=True):
        self._plans.clear()
        return super()._apply(fn, recurse)

    def _build_plan(self, device, max_batch, H, W) -> Plan:
        p = Plan(device, max_batch)
        p.in_hw = (H, W)
        bb, nk, nc = self.backbone, self.neck, self.num_classes
        c2, c3, c4 = bb.out_channels
        h2, w2 = H // 4, W // 4
        # concat buffers of the neck (model.py:257-267): producers write their slice directly
        cat_f3 = p.buffer(h2 // 2, w2 // 2, 2 * c3)   # [up(lateral_p3(sppf)), p3]
        cat_f2 = p.buffer(h2, w2, 2 * c2)             # [up(lateral_p2(f3)), p2]
        cat_o3 = p.buffer(h2 // 2, w2 // 2, c2 + c3)  # [down1(f2), f3]
        cat_o4 = p.buffer(h2 // 4, w2 // 4, c3 + c4)  # [down2(o3), p4]
        x = bb.stem.emit(p, NETWORK_INPUT)
        x = bb.stage1_conv.emit(p, x)
        p2 = bb.stage1_block.emit(p, x, cat_f2.sub(c2, c2))
        p3 = bb.stage2_c3k2.emit(p, bb.stage2_conv.emit(p, p2), cat_f3.sub(c3, c3))
        p4 = bb.stage3_c3k2.emit(p, bb.stage3_conv.emit(p, p3), cat_o4.sub(c3, c4))
        ctx = bb.sppf.emit(p, p4)
        p.upsample2x(nk.lateral_p3.emit(p, ctx), cat_f3.sub(0, c3))
        f3 = nk.fpn_c3k2_1.emit(p, cat_f3, cat_o3.sub(c2, c3))
        p.upsample2x(nk.lateral_p2.emit(p, f3), cat_f2.sub(0, c2))
        f2 = nk.fpn_c3k2_2.emit(p, cat_f2)
        nk.down1.emit(p, f2, cat_o3.sub(0, c2))
        o3 = nk.pan_c3k2_1.emit(p, cat_o3)
        nk.down2.emit(p, o3, cat_o4.sub(0, c3))
        o4 = nk.pan_c3k2_2.emit(p, cat_o4)
        heads = []
        for lvl, f in zip((2, 3, 4), (f2, o3, o4)):
            head = p.buffer(f.h, f.w, nc + 4, UYD_F32)
            getattr(self, f"head_p{lvl}").emit(p, f, head, nc)
            heads.append(head)
        p.heads = heads
        p.feature_slices = {"p2": p2, "p3": p3, "p4": p4, "sppf": ctx, "f2": f2, "o3": o3, "o4": o4}
        return p.finalize()

    def plan_for(self, x) -> Plan:
        B, _, H, W = x.shape
        dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
        p = self._plans.get((dev, H, W))
        if p is None or p.max_batch < B:
            if H % 32 or W % 32:
                raise ValueError("frame height/width must be multiples of 32")
            p = self._build_plan(dev, B, H, W)
            self._plans[(dev, H, W)] = p
        return p

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        """``[(cls, reg)] x 3`` NCHW fp32 (ONNX names p2_cls, p2_reg, ... model.py:382-383)."""
        if self.training:
            raise RuntimeError("UninaCustomB200 implements the inference path only: call .eval()")
        if not torch.cuda.is_available():
            raise _lib.UydError("no CUDA device: the B200 path has no CPU fallback")
        if not x.is_cuda:
            x = x.cuda(non_blocking=True)
        x = x.contiguous() if x.dtype == torch.uint8 else x.float().contiguous()
        p = self.plan_for(x)
        B, nc = x.shape[0], self.num_classes
        p.run(x)
        outs = []
        for h in p.heads:
            t = torch.empty(B, nc + 4, h.h, h.w, dtype=torch.float32, device=x.device)
            self._export(p, h, t, B)
            outs.append((t[:, :nc], t[:, nc:]))
        return outs

    @staticmethod
    def _export(p: Plan, h: Slice, out: torch.Tensor, batch: int):
        # NHWC fp32 head buffer -> NCHW through a plain permute of a device copy
        out.copy_(p.read(h, batch))

    @torch.no_grad()
    def predict(self, x: torch.Tensor, conf: float = 0.5, iou: float = 0.45, conformal_q: float = 0.0,
                strict: bool = True, cap: int = 0):
        """Decode + NMS with the reference's runtime thresholds (config/params.yaml:13-14).
        Returns per-image ``[N,6]`` tensors (x1,y1,x2,y2,conf,cls), N <= 1024 (MAX_DETECTIONS)."""
        outs = self.forward(x)
        B = outs[0][0].shape[0]
        dev = outs[0][0].device
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        L = _lib.lib()
        ctx = _lib.context(di)
        cells = sum(c.shape[2] * c.shape[3] for c, _ in outs)
        cap = cap or cells
        ws_bytes = int(L.uyd_nms_detections_workspace_bytes(cap))
        dets = torch.zeros(cap, 8, dtype=torch.float32, device=dev)          # 32-byte records
        kept = torch.zeros(1024, 8, dtype=torch.float32, device=dev)
        cell = torch.zeros(cap, dtype=torch.int32, device=dev)
        cnt = torch.zeros(2, dtype=torch.int32, device=dev)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(di).cuda_stream)
        results = []
        for b in range(B):
            cnt.zero_()
            base = 0
            for (cls, reg), stride in zip(outs, self.STRIDES):
                gh, gw = cls.shape[2], cls.shape[3]
                cb, rb = cls[b].contiguous(), reg[b].contiguous()
                check(L.uyd_decode_tlbr(ctx, C.c_void_p(cb.data_ptr()), C.c_void_p(rb.data_ptr()), C.c_void_p(dets.data_ptr()),
                                        C.c_void_p(cell.data_ptr()), C.c_void_p(cnt.data_ptr()), cap, gw, gh, stride,
                                        self.num_classes, conf, conformal_q, int(strict), base, stream), "uyd_decode_tlbr")
                base += gh * gw
            check(L.uyd_nms_detections(ctx, C.c_void_p(dets.data_ptr()), C.c_void_p(cell.data_ptr()), C.c_void_p(cnt.data_ptr()), cap, iou,
                                       C.c_void_p(ws.data_ptr()), ws_bytes, C.c_void_p(kept.data_ptr()),
                                       C.c_void_p(cnt[1:].data_ptr()), stream), "uyd_nms_detections")
            n = int(cnt[1].item())
            k = kept[:n]
            rows = torch.cat((k[:, :5], k[:, 5:6].view(torch.int32).float()), 1)
            results.append(rows.clone())
        return results
