"""Oracle: integer reference of the INT8 (QAT fake-quant) convolution.  TEST INFRASTRUCTURE.

PARITY UNPINNED: ``pytorch_quantization`` (demo.ipynb cell 1 pins 2.1.2; unpinned in
setup_env.sh:25) is absent, so the fake-quant arithmetic is restated from its published
behaviour as configured by the reference (qat.py:109-124: 8 bit, per-tensor ``axis=None`` for
inputs *and* weights, narrow range; qat.py:700-753: layers ``model.{0,1,2}`` stay float):

    scale = 127 / amax                    (fp32)
    q     = clamp(round_half_even(x * scale), -127, 127)
    fake-quant output = q / scale

Integer reference of one quantised Conv+BN+ReLU:

    acc = sum q_x * q_w                   (int32, exact)
    y   = float32(acc) * m_c + b_c        (separate fp32 multiply then add, round-to-nearest)
          m_c = (amax_x/127) * (amax_w/127) * gamma_c / sqrt(var_c + eps)
          b_c = beta_c - mu_c * gamma_c / sqrt(var_c + eps)
    y   = relu(y)
    q_y = clamp(round_half_even(y * (127 / amax_next)), -127, 127)     (the consumer's input quantiser)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def scale_of(amax: float) -> np.float32:
    return np.float32(127.0) / np.float32(amax)


def quantize(x: np.ndarray, amax: float) -> np.ndarray:
    """fp32 tensor -> int8 with the per-tensor fake-quant scale (multiply in fp32, half-even)."""
    s = scale_of(amax)
    return np.clip(np.rint(x.astype(np.float32) * s), -127, 127).astype(np.int8)


def fold_multiplier(amax_x: float, amax_w: float, gamma=None, beta=None, mean=None, var=None, eps=1e-3, conv_bias=None):
    """Per-channel (m_c, b_c) of the requant epilogue, fp32."""
    sx = np.float32(amax_x) / np.float32(127.0)
    sw = np.float32(amax_w) / np.float32(127.0)
    if gamma is None:
        n = len(conv_bias)
        return np.full(n, sx * sw, np.float32), np.asarray(conv_bias, np.float32)
    g = (np.asarray(gamma, np.float64) / np.sqrt(np.asarray(var, np.float64) + eps))
    m = (np.float64(sx) * np.float64(sw) * g).astype(np.float32)
    b = (np.asarray(beta, np.float64) - np.asarray(mean, np.float64) * g).astype(np.float32)
    return m, b


def conv_int8(qx: np.ndarray, qw: np.ndarray, mult: np.ndarray, bias: np.ndarray, stride: int = 1, relu: bool = True,
              out_scale: float | None = None, groups: int = 1):
    """qx [N,C,H,W] int8, qw [Co,Ci/g,k,k] int8 -> (acc int32 [N,Co,OH,OW], y fp32, q_y int8 or None)."""
    k = qw.shape[2]
    acc = F.conv2d(torch.from_numpy(qx.astype(np.float64)), torch.from_numpy(qw.astype(np.float64)), stride=stride,
                   padding=k // 2, groups=groups).numpy()
    acc = np.rint(acc).astype(np.int64)
    assert np.abs(acc).max() < 2 ** 31
    acc32 = acc.astype(np.int32)
    y = acc32.astype(np.float32) * mult.astype(np.float32)[None, :, None, None]     # fp32 multiply (rounded)
    y = (y + bias.astype(np.float32)[None, :, None, None]).astype(np.float32)       # fp32 add (rounded)
    if relu:
        y = np.maximum(y, np.float32(0))
    qy = None
    if out_scale is not None:
        qy = np.clip(np.rint(y * np.float32(out_scale)), -127, 127).astype(np.int8)
    return acc32, y, qy
