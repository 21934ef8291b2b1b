"""Per-op CUDA-event table of the INT8 plan (batch 64 by default)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import unina_yolo_dla_b200 as uyd  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(args[0]) if args else 64
m = uyd.UninaYoloB200.from_yaml().init_synthetic(0).cuda()
x = torch.rand(B, 3, 640, 640, device="cuda")
m.calibrate_int8(x[:8])
p = m.plan_for(x)
m(x)
if "--ncu" in sys.argv:   # ncu --profile-from-start off: one run of the INT8 plan
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    p.run(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
ms = p.profile(x)
if "--all" in sys.argv:   # plan order
    for i, t in enumerate(ms):
        print(f"  {i:3d} {t:7.4f} ms  {p.op_info(i)[0]}")
agg = {}
for i, t in enumerate(ms):
    txt = p.op_info(i)[0]
    kind = " ".join(w for w in txt.split() if not w[0].isdigit() and "x" not in w[1:4])[:40]
    agg.setdefault(txt.split()[0] + (" tc" if "tc:" in txt else " direct" if "direct" in txt else ""), [0, 0.0])
    a = agg[txt.split()[0] + (" tc" if "tc:" in txt else " direct" if "direct" in txt else "")]
    a[0] += 1
    a[1] += t
print(f"INT8 plan batch {B}: {len(ms)} ops, {sum(ms):.3f} ms")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:<20} x{n:3d}  {t:8.3f} ms")
for i in sorted(range(len(ms)), key=lambda i: -ms[i])[:25]:
    print(f"  {ms[i]:7.4f} ms  {p.op_info(i)[0]}")
print("quantize ops:")
for i in sorted((i for i in range(len(ms)) if p.op_info(i)[0].startswith("quantize")), key=lambda i: -ms[i])[:30]:
    print(f"  {ms[i]:7.4f} ms  {p.op_info(i)[0]}")
