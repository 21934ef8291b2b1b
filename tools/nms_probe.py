"""Candidate statistics of the bench workload and NMS timing vs candidate count (GPU box)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import unina_yolo_dla_b200 as uyd  # noqa: E402
import uyd_testlib_cpu as T  # noqa: E402


def t_ms(fn, reps=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


m = uyd.UninaYoloB200.from_yaml().init_synthetic(0).cuda()
x = torch.rand(64, 3, 640, 640, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
m.calibrate_cls_bias(x[:8], 1500, 0.25)
y = m.forward(x, raw_heads=False)
c = (y[:, 4:].amax(1) > 0.25).sum(1)
print("bench candidates/img: min", int(c.min()), "mean", float(c.float().mean()), "max", int(c.max()))
det, cnt = m.nms(y, 0.25, 0.7, 300)
print("kept/img: min", int(cnt.min()), "mean", float(cnt.float().mean()), "  nms ms", t_ms(lambda: m.nms(y, 0.25, 0.7, 300)))
for frac, cluster in ((0.02, False), (0.05, False), (0.05, True), (0.2, True), (0.9, True)):
    ys = torch.from_numpy(T.synth_predictions(64, 4, 33600, seed=1, frac_conf=frac, cluster=cluster)).cuda()
    det, cnt = m.nms(ys, 0.25, 0.7, 300)
    n = (ys[:, 4:].amax(1) > 0.25).sum(1).float().mean()
    print(f"synthetic frac {frac} cluster {cluster}: cand/img {float(n):.0f} kept/img {float(cnt.float().mean()):.0f} nms ms {t_ms(lambda: m.nms(ys, 0.25, 0.7, 300)):.3f}")
# phase split on the bench prediction: a threshold nothing passes = the scan alone; max_det = 1 = scan + sort + one chunk
print("scan only (conf 2.0): ms", t_ms(lambda: m.nms(y, 2.0, 0.7, 300)))
print("scan + sort + first chunk (max_det 1): ms", t_ms(lambda: m.nms(y, 0.25, 0.7, 1)))
print("iou 1.0 (nothing suppressed, stops at 300 kept): ms", t_ms(lambda: m.nms(y, 0.25, 1.0, 300)))
