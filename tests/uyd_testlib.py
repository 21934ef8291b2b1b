"""Helpers shared by the GPU parity tests and tools/gpu_probe.py."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

import unina_yolo_dla_b200 as uyd
from unina_yolo_dla_b200._lib import IMPL_AUTO, IMPL_DIRECT, IMPL_TC, UYD_BF16, UYD_F32  # noqa: F401


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).float()


def run_single_conv(cin, cout, k, stride, H, W, batch=2, impl=IMPL_AUTO, relu=True, residual=False, depthwise=False,
                    in_total=None, in_coff=0, out_total=None, out_coff=0, out_f32=False, seed=0, max_batch=None):
    """Builds a one-conv plan, runs it, returns (got NCHW fp32, torch fp32 reference on the
    same bf16-rounded operands)."""
    g = torch.Generator().manual_seed(seed)
    in_total = in_total or cin
    out_total = out_total or cout
    x = torch.randn(batch, cin, H, W, generator=g)
    wshape = (cout, 1, k, k) if depthwise else (cout, cin, k, k)
    w = torch.randn(*wshape, generator=g) / (wshape[1] * k * k) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    oh, ow = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
    r = torch.randn(batch, cout, oh, ow, generator=g) if residual else None

    p = uyd.Plan(0, max_batch or batch)
    src = p.buffer(H, W, in_total)
    dst = p.buffer(oh, ow, out_total, UYD_F32 if out_f32 else UYD_BF16)
    rs = p.buffer(oh, ow, cout) if residual else None
    s_in, s_out = src.sub(in_coff, cin), dst.sub(out_coff, cout)
    p.conv(s_in, s_out, w.numpy(), b.numpy(), k, stride, relu=relu, depthwise=depthwise, res=rs, impl=impl)
    p.finalize()
    p.write(s_in, x)
    if residual:
        p.write(rs, r)
    p.run_no_input(batch)
    torch.cuda.synchronize()
    got = p.read(s_out, batch).cpu()
    ref = F.conv2d(bf16_round(x), bf16_round(w), b, stride=stride, padding=k // 2, groups=cin if depthwise else 1)
    if relu:
        ref = ref.relu()
    if residual:
        ref = ref + bf16_round(r)
    if not out_f32:
        ref = bf16_round(ref)
    # untouched channels of a wider output buffer must stay zero
    if out_total != cout:
        whole = p.read(dst, batch).cpu()
        mask = torch.ones(out_total, dtype=torch.bool)
        mask[out_coff:out_coff + cout] = False
        assert float(whole[:, mask].abs().max()) == 0.0, "conv wrote outside its output slice"
    return got, ref


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12))


def synth_predictions(B, nc, A, seed=0, frac_conf=0.2, img=640.0, cluster=False):
    """Synthetic y[B,4+nc,A] with overlapping boxes and a controllable share above conf."""
    rng = np.random.default_rng(seed)
    y = np.zeros((B, 4 + nc, A), np.float32)
    if cluster:
        centers = rng.uniform(40, img - 40, (B, 2, 64))
        pick = rng.integers(0, 64, (B, A))
        cxy = np.take_along_axis(centers, pick[:, None, :].repeat(2, 1), 2) + rng.normal(0, 6, (B, 2, A))
    else:
        cxy = rng.uniform(0, img, (B, 2, A))
    y[:, 0:2] = cxy
    y[:, 2:4] = rng.uniform(8, 96, (B, 2, A))
    sc = rng.uniform(0, 1, (B, nc, A)).astype(np.float32)
    lift = rng.uniform(0, 1, (B, 1, A)) < frac_conf
    y[:, 4:] = np.where(lift, 0.25 + 0.75 * sc, 0.2 * sc)
    return y
