"""Runs a handful of single-conv plans at bench size (for ncu captures and quick timing).
Usage: python tools/conv_cases.py [--batch 64] [--reps 3] [--only halo64,flat64]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import unina_yolo_dla_b200 as uyd  # noqa: E402
from unina_yolo_dla_b200._lib import IMPL_AUTO, IMPL_DIRECT, IMPL_TC, UYD_BF16, UYD_F32  # noqa: E402

CASES = {
    # name: (cin, cout, k, stride, H, W, out_f32, depthwise, impl)
    "halo64": (64, 64, 3, 1, 160, 160, False, False, IMPL_TC),
    "halo32": (32, 64, 3, 1, 160, 160, False, False, IMPL_TC),
    "flat64": (64, 64, 1, 1, 160, 160, False, False, IMPL_TC),
    "flat64f32": (64, 64, 1, 1, 160, 160, True, False, IMPL_TC),
    "flat96_16": (96, 16, 1, 1, 160, 160, False, False, IMPL_TC),
    "flat32": (32, 32, 1, 1, 160, 160, False, False, IMPL_TC),
    "s2_16_32": (16, 32, 3, 2, 320, 320, False, False, IMPL_TC),
    "dw32": (32, 32, 3, 1, 160, 160, False, True, IMPL_DIRECT),
    "c4": (4, 4, 3, 1, 160, 160, False, False, IMPL_DIRECT),
    "c8": (8, 8, 3, 1, 80, 80, False, False, IMPL_DIRECT),
    "halo16": (16, 16, 3, 1, 40, 40, False, False, IMPL_TC),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    names = [n for n in a.only.split(",") if n] or list(CASES)
    g = torch.Generator().manual_seed(0)
    for name in names:
        cin, cout, k, s, H, W, f32, dw, impl = CASES[name]
        oh, ow = H // s, W // s
        p = uyd.Plan(0, a.batch)
        src = p.buffer(H, W, cin)
        dst = p.buffer(oh, ow, cout, UYD_F32 if f32 else UYD_BF16)
        w = torch.randn(cout, 1 if dw else cin, k, k, generator=g) * 0.05
        p.conv(src, dst, w.numpy(), torch.zeros(cout).numpy(), k, s, relu=True, depthwise=dw, impl=impl)
        p.finalize()
        p.write(src, torch.randn(a.batch, cin, H, W, generator=g))
        p.run_no_input(a.batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            p.run_no_input(a.batch)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        text, fl, by = p.op_info(0)
        print(f"{name:10s} {text:45s} {ms:8.4f} ms  {fl * a.batch / ms / 1e9:8.1f} TFLOP/s  {by * a.batch / ms / 1e6:8.0f} GB/s", flush=True)
        del p


if __name__ == "__main__":
    main()
