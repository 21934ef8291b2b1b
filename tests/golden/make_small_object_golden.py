"""Generates tests/golden/small_object_metric.npz by running the REAL reference class
``data_loader.SmallObjectMetric`` (imported from /root/reference/unina_yolo_dla) on seeded cases.
Run from the repo root:  python tests/golden/make_small_object_golden.py   (needs the read-only mount)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, "/root/reference/unina_yolo_dla")
OUT = Path(__file__).resolve().parent / "small_object_metric.npz"


def cases(seed=0, n_img=24, max_gt=12, max_pred=40):
    """Padded arrays: pred [I, P, 6], npred [I], gt [I, G, 5], ngt [I].  Predictions = jittered copies of ground truths
    (so that IoUs land on both sides of the threshold) + clutter; distinct confidences; some images without small
    ground truths, some without predictions, some with duplicated predictions (one-to-one matching matters)."""
    rng = np.random.default_rng(seed)
    pred = np.zeros((n_img, max_pred, 6), np.float32)
    gt = np.zeros((n_img, max_gt, 5), np.float32)
    npred, ngt = np.zeros(n_img, np.int32), np.zeros(n_img, np.int32)
    for i in range(n_img):
        g = int(rng.integers(0, max_gt + 1))
        wh = rng.uniform(4, 20, (g, 2)) / 640 if i % 5 else rng.uniform(20, 60, (g, 2)) / 640   # every fifth image: no small boxes
        xy = rng.uniform(0.1, 0.9, (g, 2))
        gt[i, :g] = np.concatenate((rng.integers(0, 4, (g, 1)), xy, wh), 1)
        ngt[i] = g
        rows = []
        if i % 7 != 3:
            for j in range(g):
                for _ in range(int(rng.integers(0, 3))):       # 0, 1 or 2 predictions per ground truth
                    jit = rng.normal(0, 0.15, 2) * gt[i, j, 3:5]
                    sc = rng.uniform(0.8, 1.25, 2)
                    cls = gt[i, j, 0] if rng.uniform() < 0.85 else (gt[i, j, 0] + 1) % 4
                    rows.append([gt[i, j, 1] + jit[0], gt[i, j, 2] + jit[1], gt[i, j, 3] * sc[0], gt[i, j, 4] * sc[1], 0, cls])
            for _ in range(int(rng.integers(0, 8))):            # clutter, small and large
                rows.append([*rng.uniform(0.1, 0.9, 2), *(rng.uniform(4, 30, 2) / 640), 0, rng.integers(0, 4)])
        rows = rows[:max_pred]
        if rows:
            r = np.asarray(rows, np.float32)
            r[:, 4] = rng.permutation(len(r)).astype(np.float32) / len(r) * 0.7 + 0.25    # distinct confidences
            pred[i, : len(r)] = r
        npred[i] = len(rows)
    return pred, npred, gt, ngt


def main():
    from data_loader import SmallObjectMetric  # the real reference class

    pred, npred, gt, ngt = cases()
    out = {}
    for thr in (0.5, 0.3):
        m = SmallObjectMetric(size_threshold=15, iou_threshold=thr, image_size=640)
        running = []
        for i in range(len(pred)):
            m.update([torch.from_numpy(pred[i, : npred[i]])], [torch.from_numpy(gt[i, : ngt[i]])])
            running.append((m.true_positives, m.false_positives, m.false_negatives))
        out[f"running_thr{thr}"] = np.asarray(running, np.int64)
        c = m.compute()
        out[f"prf_thr{thr}"] = np.asarray([c["small_object_precision"], c["small_object_recall"], c["small_object_f1"]])
    np.savez_compressed(OUT, pred=pred, npred=npred, gt=gt, ngt=ngt, **out)
    print(OUT, {k: v[-1].tolist() if v.ndim > 1 else v.tolist() for k, v in out.items()})


if __name__ == "__main__":
    main()
