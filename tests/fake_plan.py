"""A torch-fp32 stand-in for the libuyd plan, used ONLY by CPU tests to check the host-side
graph emission (slices, in-place concat, residual wiring).  It has the same construction
interface as unina_yolo_dla_b200.plan.Plan; nothing in the product imports it."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from unina_yolo_dla_b200.plan import Slice


class FakePlan:
    def __init__(self, device=0, max_batch=1):
        self.max_batch = max_batch
        self.specs = []       # (h, w, c)
        self.shapes = {}
        self.ops = []
        self.heads = []
        self.in_hw = None

    def buffer(self, h, w, c, dtype=0):
        self.specs.append((h, w, c))
        self.shapes[len(self.specs) - 1] = (h, w, c, dtype)
        return Slice(len(self.specs) - 1, 0, c, h, w)

    def conv(self, src, dst, weight, bias, k, stride=1, relu=True, depthwise=False, res=None, impl=0, pre=None):
        self.ops.append(("conv", src, dst, torch.as_tensor(weight).float(), torch.as_tensor(bias).float(), k, stride, relu, depthwise, res, pre))
        return dst

    @staticmethod
    def c3k_supported(src, dst):
        return src.c in (8, 16, 32) and src.w % 40 == 0 and (src.h % 32 == 0 or src.h % 20 == 0 or src.h % 16 == 0)

    def c3k(self, src, dst, weights, biases):
        self.ops.append(("c3k", src, dst, [torch.from_numpy(w) for w in weights], [torch.from_numpy(b) for b in biases]))
        return dst

    @staticmethod
    def cls_branch_supported(src, mid, nc):
        return src.c in (32, 64) and mid == 32 and nc <= 8 and src.w % 40 == 0 and src.h % 8 == 0

    def cls_branch(self, src, dst, mid, weights, biases):
        self.ops.append(("cls", src, dst, [torch.from_numpy(w) for w in weights], [torch.from_numpy(b) for b in biases]))
        return dst

    def stem2(self, dst, w0, b0, w1, b1, w2=None, b2=None):
        t = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float32)
        self.ops.append(("stem2", dst, t(w0), t(b0), t(w1), t(b1), t(w2), t(b2)))
        return dst

    @staticmethod
    def chain_supported(src, n1, n2):
        return src.c in (32, 64) and n1 in (32, 64) and 1 <= n2 <= 64 and src.coff % 8 == 0

    def chain(self, src, w1, b1, w2, b2, *, dw1=False, relu2=False, final=0, out=None, w3=None, b3=None, a_total=0, a_off=0,
              y_ch0=0, no=0, stride=1.0):
        t = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float32)
        self.ops.append(("chain", src, out, t(w1), t(b1), t(w2), t(b2), dw1, relu2, final, t(w3), t(b3),
                         (a_total, a_off, y_ch0, no, stride)))
        return out

    def sppf_pool(self, s, c):
        self.ops.append(("sppf", s, c))

    def upsample2x(self, src, dst):
        self.ops.append(("up", src, dst))
        return dst

    def set_heads(self, heads, strides, reg_max, nc):
        self.heads = heads

    def finalize(self):
        return self

    def execute(self, x):
        B = x.shape[0]
        bufs = [torch.zeros(B, c, h, w) for (h, w, c) in self.specs]

        def get(s):
            return x if s.buf < 0 else bufs[s.buf][:, s.coff:s.coff + s.c]

        self.y = None
        for op in self.ops:
            if op[0] == "stem2":
                _, dst, w0, b0, w1, b1, w2, b2 = op
                t = F.conv2d(x, w0, b0, stride=2, padding=1).relu()
                t = F.conv2d(t, w1, b1, stride=2, padding=1).relu()
                if w2 is not None:
                    t = F.conv2d(t, w2.reshape(16, 32, 1, 1), b2).relu()
                bufs[dst.buf][:, dst.coff:dst.coff + dst.c] = t
            elif op[0] == "chain":
                _, src, out, w1, b1, w2, b2, dw1, relu2, final, w3, b3, (a_total, a_off, y_ch0, no, stride) = op
                xin = get(src)
                t = F.conv2d(xin, w1, b1, padding=1, groups=xin.shape[1] if dw1 else 1).relu()
                u = F.conv2d(t, w2.reshape(w2.shape[0], -1, 1, 1), b2)
                hw = src.h * src.w
                if self.y is None and a_total and final != 0 and out is None:
                    self.y = torch.zeros(B, no, a_total)
                if final == 0:  # STORE
                    bufs[out.buf][:, out.coff:out.coff + out.c] = u.relu() if relu2 else u
                elif final == 2:  # PW3
                    lg = F.conv2d(u.relu(), w3.reshape(w3.shape[0], -1, 1, 1), b3)
                    if out is not None:
                        bufs[out.buf][:, out.coff:out.coff + out.c] = lg
                    else:
                        self.y[:, y_ch0:y_ch0 + lg.shape[1], a_off:a_off + hw] = lg.sigmoid().flatten(2)
                else:  # DFL
                    if out is not None:
                        bufs[out.buf][:, out.coff:out.coff + out.c] = u
                    else:
                        d = (u.view(B, 4, 16, hw).softmax(2) * torch.arange(16.0).view(1, 1, 16, 1)).sum(2)  # l, t, r, b
                        ax = (torch.arange(hw) % src.w).float() + 0.5
                        ay = (torch.arange(hw) // src.w).float() + 0.5
                        x1, y1, x2, y2 = ax - d[:, 0], ay - d[:, 1], ax + d[:, 2], ay + d[:, 3]
                        self.y[:, y_ch0:y_ch0 + 4, a_off:a_off + hw] = torch.stack(((x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1), 1) * stride
            elif op[0] == "conv":
                _, src, dst, w, b, k, stride, relu, dw, res, pre = op
                y = F.conv2d(get(src), w, b, stride=stride, padding=k // 2, groups=w.shape[0] if dw else 1)
                if pre is not None:
                    y = y + F.interpolate(get(pre), scale_factor=2, mode="nearest")
                if relu:
                    y = y.relu()
                if res is not None:
                    y = y + get(res)
                bufs[dst.buf][:, dst.coff:dst.coff + dst.c] = y
            elif op[0] == "cls":
                _, src, dst, w, b = op
                t = get(src)
                t = F.conv2d(t, w[0], b[0], padding=1, groups=w[0].shape[0]).relu()
                t = F.conv2d(t, w[1], b[1]).relu()
                t = F.conv2d(t, w[2], b[2], padding=1, groups=w[2].shape[0]).relu()
                t = F.conv2d(t, w[3], b[3]).relu()
                bufs[dst.buf][:, dst.coff:dst.coff + dst.c] = F.conv2d(t, w[4], b[4])
            elif op[0] == "c3k":
                _, src, dst, w, b = op
                cb = lambda t, i, k: F.conv2d(t, w[i], b[i], padding=k // 2).relu()
                xin = get(src)
                a_, b_ = cb(xin, 0, 1), cb(xin, 1, 1)
                u = a_ + cb(cb(a_, 2, 3), 3, 3)
                v = u + cb(cb(u, 4, 3), 5, 3)
                bufs[dst.buf][:, dst.coff:dst.coff + dst.c] = cb(torch.cat((v, b_), 1), 6, 1)
            elif op[0] == "sppf":
                _, s, c = op
                t = get(s.sub(0, c))
                for i in range(1, 4):
                    t = F.max_pool2d(t, 5, 1, 2)
                    bufs[s.buf][:, s.coff + i * c:s.coff + (i + 1) * c] = t
            else:
                _, src, dst = op
                bufs[dst.buf][:, dst.coff:dst.coff + dst.c] = F.interpolate(get(src), scale_factor=2, mode="nearest")
        return bufs
