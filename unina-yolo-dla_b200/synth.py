"""Seeded, data-free initialisation shared by both model facades (parity runs, benchmarks)."""
from __future__ import annotations

import torch
import torch.nn as nn


@torch.no_grad()
def synthetic_init_(module: nn.Module, seed: int = 0, gain: float = 1.0) -> nn.Module:
    """Normal conv weights with variance gain/fan_in, small biases, randomised BN affine and
    running statistics (so BN folding is exercised).  gain = 1 gives a contractive network on
    which 2^-9 bf16 rounding noise is not amplified with depth."""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, nn.Conv2d) and m.weight.requires_grad:
            fan_in = m.weight.shape[1] * m.weight.shape[2] * m.weight.shape[3]
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (gain / fan_in) ** 0.5)
            if m.bias is not None:
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.copy_(0.8 + 0.4 * torch.rand(m.weight.shape, generator=g))
            m.bias.copy_(0.2 * (torch.rand(m.bias.shape, generator=g) - 0.5))
            m.running_mean.copy_(0.2 * torch.randn(m.running_mean.shape, generator=g))
            m.running_var.copy_(0.8 + 0.4 * torch.rand(m.running_var.shape, generator=g))
    return module
