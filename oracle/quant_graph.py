"""Oracle: integer reference of the whole INT8 (QAT fake-quant) network.  TEST INFRASTRUCTURE.

PARITY UNPINNED (see oracle/quant.py): pytorch-quantization is absent; the fake-quant arithmetic is the
published one as the reference configures it (qat.py:109-124: 8 bit, narrow range, per-tensor input and
weight scales; qat.py:700-753: ``model.{0,1,2}`` stay float; train.py:725: every nn.Conv2d is wrapped,
including Detect's biased 1x1 convs and the DFL projection).

Fixed conventions of this restatement (the CUDA path implements exactly these, so raw head outputs are
compared byte for byte):
  * activations between layers are bf16 values (the storage type of the B200 path); every quantised conv
    computes  y = relu(float32(acc) * m_c + b_c) [+ residual, fp32]  and rounds y once to bf16;
  * the final biased convs of Detect produce fp32 (no rounding);
  * BN / ReLU / residual add / concat / max-pool / upsample are floating point, as in the QAT graph;
  * the float layers are NOT restated here: the comparison starts from the tensor they produce.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import quant as oq
from . import yolo_graph as yg


def bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).float()


class Int8Graph:
    def __init__(self, model: yg.DetectionModel, amax: dict, float_layers=(0, 1, 2)):
        self.model, self.amax, self.float_layers = model, amax, set(float_layers)

    # ---- one QuantConv2d (+ BN + ReLU) ---------------------------------------------------------
    def qconv(self, conv: torch.nn.Conv2d, bn, name: str, x: torch.Tensor, relu: bool, res=None, out_bf16=True):
        ax, aw = self.amax[name]
        qx = oq.quantize(x.numpy(), ax)
        qw = oq.quantize(conv.weight.detach().numpy(), aw)
        if bn is not None:
            mult, bias = oq.fold_multiplier(ax, aw, bn.weight.detach().numpy(), bn.bias.detach().numpy(),
                                            bn.running_mean.numpy(), bn.running_var.numpy(), bn.eps)
        else:
            mult, bias = oq.fold_multiplier(ax, aw, conv_bias=conv.bias.detach().numpy())
        _, y, _ = oq.conv_int8(qx, qw, mult, bias, conv.stride[0], relu=relu, groups=conv.groups)
        y = torch.from_numpy(y)
        if res is not None:
            y = y + res                      # fp32 add after the activation (Bottleneck shortcut)
        return bf16(y) if out_bf16 else y

    def conv(self, m: yg.Conv, name: str, x, res=None):
        return self.qconv(m.conv, m.bn, name + ".conv", x, True, res)

    # ---- module walk (mirrors oracle/yolo_graph.py forwards) -------------------------------------
    def run(self, m, name: str, x):
        if isinstance(m, yg.Conv):
            return self.conv(m, name, x)
        if isinstance(m, yg.Bottleneck):
            t = self.conv(m.cv1, name + ".cv1", x)
            return self.conv(m.cv2, name + ".cv2", t, res=x if m.add else None)
        if isinstance(m, yg.C3k):
            a = self.conv(m.cv1, name + ".cv1", x)
            for j, b in enumerate(m.m):
                a = self.run(b, f"{name}.m.{j}", a)
            return self.conv(m.cv3, name + ".cv3", torch.cat((a, self.conv(m.cv2, name + ".cv2", x)), 1))
        if isinstance(m, yg.C3k2):
            y = list(self.conv(m.cv1, name + ".cv1", x).chunk(2, 1))
            for j, b in enumerate(m.m):
                y.append(self.run(b, f"{name}.m.{j}", y[-1]))
            return self.conv(m.cv2, name + ".cv2", torch.cat(y, 1))
        if isinstance(m, yg.SPPF_DLA):
            t = self.conv(m.cv1, name + ".cv1", x)
            y1 = m.m(t)
            y2 = m.m(y1)
            return self.conv(m.cv2, name + ".cv2", torch.cat((t, y1, y2, m.m(y2)), 1))
        if isinstance(m, (torch.nn.Upsample, yg.Concat)):
            return m(x)
        raise TypeError(type(m))

    def detect_raw(self, det: yg.Detect, name: str, feats):
        out = []
        for i, f in enumerate(feats):
            t = self.conv(det.cv2[i][0], f"{name}.cv2.{i}.0", f)
            t = self.conv(det.cv2[i][1], f"{name}.cv2.{i}.1", t)
            box = self.qconv(det.cv2[i][2], None, f"{name}.cv2.{i}.2", t, False, out_bf16=False)
            c = f
            for j in range(2):
                c = self.conv(det.cv3[i][j][0], f"{name}.cv3.{i}.{j}.0", c)
                c = self.conv(det.cv3[i][j][1], f"{name}.cv3.{i}.{j}.1", c)
            cls = self.qconv(det.cv3[i][2], None, f"{name}.cv3.{i}.2", c, False, out_bf16=False)
            out.append(torch.cat((box, cls), 1))
        return out

    @torch.no_grad()
    def forward_from(self, saved: dict):
        """``saved``: {layer index: bf16-valued fp32 NCHW output} of the last float layer (and of any earlier
        layer the graph routes from).  Returns the raw heads [B, 64 + nc, H_l, W_l] fp32."""
        start = max(saved) + 1
        y = [saved.get(i) for i in range(start)]
        x = y[-1]
        for m in self.model.model[start:]:
            if m.f != -1:
                x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
            if isinstance(m, yg.Detect):
                return self.detect_raw(m, f"model.{m.i}", list(x))
            x = self.run(m, f"model.{m.i}", x)
            y.append(x)
        raise RuntimeError("graph has no Detect layer")

    def decode(self, det: yg.Detect, raws, dfl_name: str):
        """Detect decode with the DFL projection as a QuantConv2d (when ``dfl_name`` has an amax entry)."""
        if dfl_name not in self.amax:
            return det.decode(raws)
        b = raws[0].shape[0]
        x_cat = torch.cat([xi.reshape(b, det.no, -1) for xi in raws], 2)
        anchors, strides = (t.transpose(0, 1) for t in yg.make_anchors(raws, det.stride, 0.5))
        box, cls = x_cat.split((det.reg_max * 4, det.nc), 1)
        prob = box.view(b, 4, det.reg_max, -1).softmax(2)
        ax, aw = self.amax[dfl_name]
        qp = np.clip(np.rint(prob.numpy() * oq.scale_of(ax)), -127, 127)
        qw = np.rint(np.arange(det.reg_max, dtype=np.float32) * oq.scale_of(aw))
        dq = np.float32(np.float32(ax) / np.float32(127)) * np.float32(np.float32(aw) / np.float32(127))
        dist = torch.from_numpy(((qp * qw[None, None, :, None]).sum(2) * dq).astype(np.float32))
        dbox = yg.dist2bbox(dist, anchors.unsqueeze(0), xywh=True, dim=1) * strides
        return torch.cat((dbox, cls.sigmoid()), 1)
