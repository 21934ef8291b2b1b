"""CPU check of the second-generation fused stem's lane maps and weight packing (csrc/stem_v2.cuh) through a host
emulation of the CTA (tests/stem_emu.cpp) against torch: Conv(3,16,3,2)+ReLU -> Conv(16,32,3,2)+ReLU
(-> 1x1 Conv(32,16)+ReLU), bf16 weights, bf16 rounding of each layer's output (model.py / unina-yolo-dla-m.yaml
layers 0-2).  The GPU parity tests of the same kernel are tests/test_gpu_parity.py::test_fused_stem_*."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = Path(__file__).resolve().parent


@pytest.fixture(scope="module")
def emu():
    out = HERE / "_build" / "stem_emu.so"
    out.parent.mkdir(exist_ok=True)
    src = HERE / "stem_emu.cpp"
    csrc = HERE.parent / "unina-yolo-dla_b200" / "csrc"
    newest = max(p.stat().st_mtime for p in (src, csrc / "stem_v2.cuh", csrc / "c3k_flat.cuh"))
    if not out.exists() or out.stat().st_mtime < newest:
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include", str(src), "-o", str(out)], check=True)
    lib = C.CDLL(str(out))
    lib.stem_emu.restype = C.c_int
    lib.div255_mismatches.restype = C.c_int
    return lib


def _bf16(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("H,W,pw,u8", [(32, 64, 0, 0), (32, 64, 1, 0), (64, 96, 1, 1), (36, 72, 0, 1), (36, 72, 1, 0)])
def test_stem_v2_lane_maps_match_torch(emu, H, W, pw, u8):
    g = torch.Generator().manual_seed(H + W + pw)
    w0, w1 = torch.randn(16, 3, 3, 3, generator=g) / 27 ** 0.5, torch.randn(32, 16, 3, 3, generator=g) / 144 ** 0.5
    w2 = torch.randn(16, 32, generator=g) / 32 ** 0.5
    b0, b1, b2 = (torch.randn(n, generator=g) * 0.1 for n in (16, 32, 16))
    x = torch.rand(1, 3, H, W, generator=g)
    xin = x
    if u8:
        xin = (x * 255).to(torch.uint8).float()
        x = xin / 255
    oc = 16 if pw else 32
    y = np.full((H // 4, W // 4, oc), np.nan, np.float32)
    fp = C.POINTER(C.c_float)
    arrs = [np.ascontiguousarray(a.numpy()) for a in (w0, b0, w1, b1, w2, b2, xin[0])]
    assert emu.stem_emu(H, W, pw, u8, *[a.ctypes.data_as(fp) for a in arrs], y.ctypes.data_as(fp)) == 0
    t = _bf16(F.conv2d(x, _bf16(w0), b0, stride=2, padding=1).relu())
    want = _bf16(F.conv2d(t, _bf16(w1), b1, stride=2, padding=1).relu())
    if pw:
        want = _bf16(F.conv2d(want, _bf16(w2).reshape(16, 32, 1, 1), b2).relu())
    want = want[0].permute(1, 2, 0).numpy()
    assert np.isfinite(y).all()
    err = np.abs(y - want).max() / max(np.abs(want).max(), 1e-6)
    assert err < 6e-3, err


def test_div255_equals_ieee_quotient(emu):
    assert emu.div255_mismatches() == 0
