"""Debug: per-item clock64 timeline of CTA 0 of the tcgen05 C3k kernel.

Needs a debug build: UYD_NVCC_EXTRA=-DUYD_C3K_TIMELINE_BUILD python unina-yolo-dla_b200/build.py
"""
import os
import sys
from pathlib import Path

os.environ["UYD_C3K_TIMELINE"] = "1"
os.environ["UYD_C3K_TC"] = "1"
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import unina_yolo_dla_b200 as uyd  # noqa: E402

m = uyd.UninaYoloB200.from_yaml().init_synthetic(0).cuda()
x = torch.rand(64, 3, 640, 640, device="cuda")
y = m.forward(x, raw_heads=False)
torch.cuda.synchronize()
