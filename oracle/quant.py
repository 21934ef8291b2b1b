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


# ---- histogram / entropy calibration, loop form (pytorch-quantization 2.1.2 calib/histogram.py restated) ----------------
# PARITY UNPINNED: the library is not installable here; this follows its published algorithm statement by
# statement and pins the product's vectorised version (unina-yolo-dla_b200/quant.py).
def histogram_collect(batches, num_bins: int = 2048):
    """HistogramCalibrator.collect over a list of arrays -> (hist int64, edges float64)."""
    import numpy as np

    hist, width = None, None
    for x in batches:
        a = np.abs(np.asarray(x, np.float32)).reshape(-1)
        xmax = float(a.max())
        if hist is None:
            width = max(xmax, 1e-12) / num_bins
            n = num_bins
        else:
            n = max(len(hist), int(np.ceil(xmax / width - 1e-9)))
        idx = np.minimum((a * np.float32(1.0 / width)).astype(np.int64), n - 1)
        h = np.bincount(idx, minlength=n).astype(np.int64)
        if hist is not None:
            h[: len(hist)] += hist
        hist = h
    return hist, np.arange(len(hist) + 1, dtype=np.float64) * width


def amax_entropy_loops(hist, edges, num_bits=8, unsigned=False, stride=1, start_bin=128):
    import numpy as np
    from collections import Counter

    bins = [float(v) for v in hist]
    bins[0] = bins[1]
    total = sum(bins)
    nlev = 1 << (num_bits - 1 + int(unsigned))
    divergences = []
    for i in range(start_bin, len(bins) + 1, stride):
        new_counts = [0.0] * nlev
        space = np.linspace(0, i, num=nlev + 1)
        dig = list(np.digitize(range(i), space) - 1)
        for idx in range(i):
            if bins[idx] == 0:
                dig[idx] = -1
        for idx, d in enumerate(dig):
            if d != -1:
                new_counts[d] += bins[idx]
        for key, val in Counter(dig).items():
            if key != -1:
                new_counts[key] = new_counts[key] / val
        new = [new_counts[d] if d != -1 else 0.0 for d in dig]
        ref = list(bins[:i])
        ref[-1] += sum(bins[i:])
        assert round(sum(new) + sum(bins[i:])) == round(total) and round(sum(ref)) == round(total)
        sn, sr = sum(new), sum(ref)
        ent = 0.0
        for p, q in zip(ref, new):
            if p > 0:
                if q == 0:
                    ent = float("inf")
                    break
                ent += (p / sr) * np.log((p / sr) / (q / sn))
        divergences.append(ent)
    divergences = np.array(divergences)
    last_argmin = len(divergences) - 1 - int(np.argmin(divergences[::-1]))
    return float(edges[last_argmin * stride + start_bin])
