"""Layer-by-layer comparison of the CUDA plan against the CPU oracle (debug aid, GPU box).
Usage: python tools/layer_diff.py [--size 640] [--batch 1]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))

import unina_yolo_dla_b200 as uyd  # noqa: E402
from oracle import init as oi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--init", default="synthetic", choices=["synthetic", "bncal"])
    ap.add_argument("--gain", type=float, default=1.0)
    a = ap.parse_args()
    from oracle import yolo_graph as yg

    if a.init == "bncal":
        ref = oi.build_yolo(seed=0, cls_bias=-2.0)
        m = uyd.UninaYoloB200.from_yaml()
        m.load_state_dict(ref.state_dict(), strict=True)
    else:
        m = uyd.UninaYoloB200.from_yaml().init_synthetic(0, gain=a.gain)
        ref = yg.DetectionModel(yg.default_yaml_path())
        ref.load_state_dict(m.state_dict(), strict=True)
        ref.eval()
    m = m.cuda()
    x = oi.seeded_frames(a.batch, a.size, seed=5)
    with torch.no_grad():
        (y_ref, raw_ref), feats = ref.forward_features(x, want=set(range(20)))
    y, raws = m(x.cuda())
    p = m.plan_for(x.cuda())
    for i, s in enumerate(p.layer_outputs):
        if s is None or i not in feats:
            continue
        got = p.read(s, a.batch).cpu()
        want = feats[i]
        err = float((got - want).abs().max() / want.abs().max().clamp_min(1e-9))
        print(f"layer {i:2d} {type(m.model[i]).__name__:10s} shape {tuple(want.shape)} max|ref| {float(want.abs().max()):8.3f} rel_err {err:.4f}")
    for lvl, (g, w) in enumerate(zip(raws, raw_ref)):
        g = g.cpu()
        print(f"head {lvl} box rel_err {float((g[:, :64] - w[:, :64]).abs().max() / w[:, :64].abs().max()):.4f} "
              f"cls rel_err {float((g[:, 64:] - w[:, 64:]).abs().max() / w[:, 64:].abs().max()):.4f}")
    yc = y.cpu()
    print("decoded box rel_err", float((yc[:, :4] - y_ref[:, :4]).abs().max() / y_ref[:, :4].abs().max()),
          "cls abs err", float((yc[:, 4:] - y_ref[:, 4:]).abs().max()))


if __name__ == "__main__":
    main()
