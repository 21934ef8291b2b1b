"""Per-kernel CUDA time of one predict_batched step (torch.profiler / CUPTI sees libuyd's launches).
Usage: python tools/kernel_table.py [--batch 64] [--size 640] [--top 40]"""
import argparse
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import unina_yolo_dla_b200 as uyd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--cand", type=int, default=1500)
    ap.add_argument("--ncu", action="store_true", help="bracket ONE step with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    ap.add_argument("--custom", action="store_true", help="the model.py network (UninaCustomB200, base_channels 32) instead of the YAML one")
    a = ap.parse_args()
    if a.custom:
        c = uyd.UninaCustomB200(4, 32).init_synthetic(0).cuda()
        xc = torch.rand(a.batch, 3, a.size, a.size, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        plan = c.plan_for(xc)
        for _ in range(3):
            plan.run(xc)
        torch.cuda.synchronize()
        if a.ncu:
            torch.cuda.profiler.start()
            plan.run(xc)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            return
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                plan.run(xc)
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
        print(f"total CUDA time per step: {sum(e.device_time_total for e in rows) / 3 / 1e3:.3f} ms")
        for e in rows[: a.top]:
            print(f"{e.device_time_total / 3:10.1f} us  x{e.count // 3:3d}  {e.key[:110]}")
        return
    m = uyd.UninaYoloB200.from_yaml().init_synthetic(0).cuda()
    x = torch.rand(a.batch, 3, a.size, a.size, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    m.calibrate_cls_bias(x[:8], a.cand, 0.25)
    for _ in range(3):
        m.predict_batched(x)
    torch.cuda.synchronize()
    if a.ncu:
        torch.cuda.profiler.start()
        m.predict_batched(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            m.predict_batched(x)
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in rows)
    print(f"total CUDA time per step: {tot / 3 / 1e3:.3f} ms")
    for e in rows[: a.top]:
        print(f"{e.device_time_total / 3:10.1f} us  x{e.count // 3:3d}  {e.key[:110]}")


if __name__ == "__main__":
    main()
