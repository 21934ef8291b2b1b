"""Seeded, BN-calibrated random initialisation for parity runs.  TEST INFRASTRUCTURE.

Why (SURVEY.md header fact 3 / section 8d): with torch-default conv init and eval-mode BN
(mean 0 / var 1) activations decay ~10x per stage and every class score equals the head
bias, so parity on raw random weights is vacuous.  One train-mode pass with cumulative
BN statistics over seeded frames restores O(1) activations at every depth without
touching a single state_dict key.
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch
import torch.nn as nn


def seeded_frames(batch: int, size: int = 640, seed: int = 0) -> torch.Tensor:
    """Synthetic frames in [0,1): smooth low-frequency content plus noise, so that conv
    responses vary across the image (pure white noise gives nearly constant heads)."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(batch, 3, size // 16, size // 16, generator=g)
    img = torch.nn.functional.interpolate(low, size=(size, size), mode="bilinear", align_corners=False)
    img = 0.7 * img + 0.3 * torch.rand(batch, 3, size, size, generator=g)
    return img.clamp_(0, 1 - 1e-6).contiguous()


@torch.no_grad()
def calibrate_bn(model: nn.Module, frames: torch.Tensor) -> nn.Module:
    bns = [m for m in model.modules() if isinstance(m, nn.BatchNorm2d)]
    saved = [m.momentum for m in bns]
    for m in bns:
        m.reset_running_stats()
        m.momentum = None
    model.train()
    model(frames)
    model.eval()
    for m, mom in zip(bns, saved):
        m.momentum = mom
    return model


@torch.no_grad()
def perturb_bn_affine(model: nn.Module, seed: int) -> None:
    """Give gamma/beta non-trivial seeded values so that BN folding is actually exercised."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.weight.copy_(0.8 + 0.4 * torch.rand(m.weight.shape, generator=g))
            m.bias.copy_(0.2 * (torch.rand(m.bias.shape, generator=g) - 0.5))


def build_yolo(seed: int = 0, calib_batch: int = 4, size: int = 640, cls_bias: float | None = -1.6,
               yaml_path=None):
    """Oracle DetectionModel with seeded weights, perturbed BN affine, calibrated BN stats.
    ``cls_bias`` replaces the (vacuous) Ultralytics cls bias so that a realistic number of
    anchors clears the confidence threshold."""
    from . import yolo_graph as yg

    torch.manual_seed(seed)
    model = yg.DetectionModel(yaml_path or yg.default_yaml_path())
    perturb_bn_affine(model, seed + 1)
    calibrate_bn(model, seeded_frames(calib_batch, size, seed + 2))
    if cls_bias is not None:
        with torch.no_grad():
            for seq in model.model[-1].cv3:
                seq[-1].bias.fill_(cls_bias)
    return model.eval()


def build_custom(seed: int = 0, base_channels: int = 32, calib_batch: int = 4, size: int = 640,
                 reg_bias: float = 2.0, cls=None):
    """Oracle (or real, via ``cls``) model.py network, calibrated; positive reg bias so the
    TLBR boxes do not invert (SURVEY.md 8d caveat)."""
    from . import custom_graph as cg

    torch.manual_seed(seed)
    model = (cls or cg.CustomNet)(4, base_channels)
    perturb_bn_affine(model, seed + 1)
    calibrate_bn(model, seeded_frames(calib_batch, size, seed + 2))
    with torch.no_grad():
        for h in (model.head_p2, model.head_p3, model.head_p4):
            h.reg_branch[2].bias.fill_(reg_bias)
    return model.eval()


def state_dict_sha256(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes())
    return h.hexdigest()
