"""CPU check of the flat-frame C3k kernel's lane maps and weight packing (csrc/c3k_flat.cuh) through a host
emulation of the warp (tests/c3k_emu.cpp) against the seven torch convs with bf16 rounding at the same points
(Ultralytics C3k with two 3x3 bottlenecks).  The GPU parity test of the same block is
tests/test_gpu_parity.py::test_fused_c3k_matches_torch."""
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
    out = HERE / "_build" / "c3k_emu.so"
    out.parent.mkdir(exist_ok=True)
    src = HERE / "c3k_emu.cpp"
    hdr = HERE.parent / "unina-yolo-dla_b200" / "csrc" / "c3k_flat.cuh"
    if not out.exists() or out.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include", str(src), "-o", str(out)], check=True)
    lib = C.CDLL(str(out))
    lib.c3k_emu.restype = C.c_int
    return lib


def _bf16(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("c,H,W,TH", [(8, 32, 80, 32), (8, 16, 40, 8), (16, 40, 80, 20), (16, 16, 40, 4), (32, 40, 40, 20), (32, 16, 80, 16)])
def test_flat_c3k_lane_maps_match_torch(emu, c, H, W, TH):
    g = torch.Generator().manual_seed(c + H + TH)
    h = c // 2
    shapes = [(h, c, 1), (h, c, 1), (h, h, 3), (h, h, 3), (h, h, 3), (h, h, 3), (c, c, 1)]
    ws = [torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5 for co, ci, k in shapes]
    bs = [torch.randn(co, generator=g) * 0.1 for co, _, _ in shapes]
    x = _bf16(torch.randn(1, c, H, W, generator=g))
    x_nhwc = np.ascontiguousarray(x[0].permute(1, 2, 0).numpy())
    y = np.full((H, W, c), np.nan, np.float32)
    wn = [np.ascontiguousarray(w.numpy()) for w in ws]
    bn = [np.ascontiguousarray(b.numpy()) for b in bs]
    fp = C.POINTER(C.c_float)
    wp = (fp * 7)(*[a.ctypes.data_as(fp) for a in wn])
    bp = (fp * 7)(*[a.ctypes.data_as(fp) for a in bn])
    assert emu.c3k_emu(c, H, W, TH, wp, bp, x_nhwc.ctypes.data_as(fp), y.ctypes.data_as(fp)) == 0
    cb = lambda t, i: _bf16(F.conv2d(t, _bf16(ws[i]), bs[i], padding=ws[i].shape[2] // 2).relu())
    a_, b_ = cb(x, 0), cb(x, 1)
    u = _bf16(a_ + F.conv2d(cb(a_, 2), _bf16(ws[3]), bs[3], padding=1).relu())
    v = _bf16(u + F.conv2d(cb(u, 4), _bf16(ws[5]), bs[5], padding=1).relu())
    want = cb(torch.cat((v, b_), 1), 6)[0].permute(1, 2, 0).numpy()
    assert np.isfinite(y).all()
    err = np.abs(y - want).max() / max(np.abs(want).max(), 1e-6)
    assert err < 1e-2, err
