"""INT8 execution of the QAT graph (reference: qat.py:91-126 quantiser configuration, qat.py:700-753
precision carve-out, train.py:721-725 ``quant_modules.initialize()``).

Every ``nn.Conv2d`` outside the float layers is a QuantConv2d: 8-bit, narrow range, per-tensor scales for
inputs *and* weights.  With static ``_amax`` the fake-quant convolution is an integer computation:

    q_x = clamp(rne(x * 127 / amax_x), -127, 127)        (uyd_plan_add_quantize, from the bf16 activation)
    q_w = clamp(rne(w * 127 / amax_w), -127, 127)        (here, on the host)
    acc = sum q_x * q_w                                   (int32, exact: tcgen05 kind::i8 / dp4a)
    y   = float(acc) * m_c + b_c ; ReLU ; (+ residual) ; round to bf16     (conv epilogue)
          m_c = (amax_x / 127) (amax_w / 127) gamma_c / sqrt(var_c + eps),  b_c = beta_c - mu_c gamma_c / sqrt(var_c + eps)

BN, ReLU, residual adds, concat, max-pool and upsample stay floating point (bf16 activations), exactly as
in the QAT graph.  The arithmetic of every step is fixed (fp32 round-to-nearest operations in a fixed order),
so the outputs are bit-exact with respect to the integer reference of the test-suite.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

IN_SUFFIX = "._input_quantizer._amax"
W_SUFFIX = "._weight_quantizer._amax"


@dataclass
class QuantSpec:
    amax: dict = field(default_factory=dict)          # conv module name -> (input amax, weight amax)
    float_layers: tuple = (0, 1, 2)                   # model.{i} kept in floating point (qat.py:700-753)

    def covers(self, name: str) -> bool:
        if name not in self.amax:
            return False
        parts = name.split(".")
        return not (len(parts) > 1 and parts[0] == "model" and parts[1].isdigit() and int(parts[1]) in self.float_layers)


def scale_of(amax: float) -> np.float32:
    return np.float32(127.0) / np.float32(amax)


def quantize_weights(w: torch.Tensor, amax: float) -> np.ndarray:
    """Weight quantiser: per-tensor, fp32 multiply, round half to even, narrow range."""
    s = scale_of(amax)
    return np.clip(np.rint(w.detach().float().cpu().numpy().astype(np.float32) * s), -127, 127).astype(np.int8)


def requant_params(amax_x: float, amax_w: float, bn: torch.nn.BatchNorm2d | None, conv_bias: torch.Tensor | None):
    """Per-channel (m_c, b_c) of the requant epilogue, fp32."""
    sx = np.float32(amax_x) / np.float32(127.0)
    sw = np.float32(amax_w) / np.float32(127.0)
    if bn is None:
        b = conv_bias.detach().float().cpu().numpy().astype(np.float32)
        return np.full(b.shape[0], sx * sw, np.float32), b
    g = bn.weight.detach().double().cpu().numpy() / np.sqrt(bn.running_var.detach().double().cpu().numpy() + bn.eps)
    m = (np.float64(sx) * np.float64(sw) * g).astype(np.float32)
    b = (bn.bias.detach().double().cpu().numpy() - bn.running_mean.detach().double().cpu().numpy() * g).astype(np.float32)
    return m, b


def split_state_dict(sd: dict):
    """Separates pytorch-quantization ``_amax`` buffers from the module parameters."""
    plain, amax_in, amax_w = {}, {}, {}
    for k, v in sd.items():
        if k.endswith(IN_SUFFIX):
            amax_in[k[: -len(IN_SUFFIX)]] = float(torch.as_tensor(v).max())
        elif k.endswith(W_SUFFIX):
            amax_w[k[: -len(W_SUFFIX)]] = float(torch.as_tensor(v).max())
        else:
            plain[k] = v
    amax = {n: (amax_in[n], amax_w[n]) for n in amax_in if n in amax_w}
    return plain, amax
