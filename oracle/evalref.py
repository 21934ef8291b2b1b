"""Oracle: the evaluation consumers of the [N, 6] detections, restated.  TEST INFRASTRUCTURE.

* ``small_object_counts``  follows UninaValidator.update_metrics, trainer.py:210-265.  PARITY UNPINNED: the
  function lives in a class that needs ``ultralytics`` (absent); ``box_iou`` below restates
  ultralytics.utils.metrics.box_iou (inter / (area1 + area2 - inter + 1e-7), fp32).
* ``conformal_scores`` / ``conformal_quantile`` follow calibrate_conformal_prediction, train.py:335-350
  (its nested box_iou), :440-470 (greedy matching) and :491-512 (numpy quantile).  PARITY UNPINNED: the
  function needs a trained YOLO object and a dataset on disk; the loops below are its loops.
Pure-Python loops: small cases only.
"""
from __future__ import annotations

import numpy as np
import torch


def box_iou(b1: torch.Tensor, b2: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    (a1, a2), (c1, c2) = b1.float().unsqueeze(1).chunk(2, 2), b2.float().unsqueeze(0).chunk(2, 2)
    inter = (torch.min(a2, c2) - torch.max(a1, c1)).clamp_(0).prod(2)
    return inter / ((a2 - a1).prod(2) + (c2 - c1).prod(2) - inter + eps)


def small_object_counts(preds, gts, size_threshold=15.0, iou_thr=0.45):
    """preds: list of [N_i, 6] (x1,y1,x2,y2,conf,cls); gts: list of [G_i, 5] (cls,x1,y1,x2,y2) pixels."""
    tp = fp = fn = 0
    for pred, gt in zip(preds, gts):
        if gt.numel() == 0:
            continue
        gt_xyxy, gt_cls = gt[:, 1:5], gt[:, 0]
        gw, gh = gt_xyxy[:, 2] - gt_xyxy[:, 0], gt_xyxy[:, 3] - gt_xyxy[:, 1]
        small = (gw < size_threshold) & (gh < size_threshold)
        if not small.any():
            continue
        sg, sc = gt_xyxy[small], gt_cls[small]
        n_small = int(small.sum())
        if pred.numel() == 0:
            fn += n_small
            continue
        pb, pc = pred[:, :4], pred[:, 5]
        pw, ph = pb[:, 2] - pb[:, 0], pb[:, 3] - pb[:, 1]
        sp = (pw < size_threshold) & (ph < size_threshold)
        match = (box_iou(pb, sg) > iou_thr) & (pc.view(-1, 1) == sc.view(1, -1))
        t = int(match.any(0).sum())
        tp += t
        fn += n_small - t
        fp += int(sp.sum()) - int(match[sp].any(1).sum())
    return tp, fp, fn


def _iou_plain(b1, b2):
    x1, y1, x2, y2 = max(b1[0], b2[0]), max(b1[1], b2[1]), min(b1[2], b2[2]), min(b1[3], b2[3])
    if x2 <= x1 or y2 <= y1:
        return np.float32(0.0)
    inter = (x2 - x1) * (y2 - y1)
    union = (b1[2] - b1[0]) * (b1[3] - b1[1]) + (b2[2] - b2[0]) * (b2[3] - b2[1]) - inter
    return inter / union if union > 0 else np.float32(0.0)


def conformal_scores(preds, gts, match_iou=0.5):
    """Nonconformity scores 1 - IoU of the greedily matched pairs, fp32 arithmetic, in processing order."""
    out = []
    for pred, gt in zip(preds, gts):
        if pred.numel() == 0:
            continue
        p = pred.numpy().astype(np.float32)
        g = gt.numpy().astype(np.float32)
        matched = set()
        for idx in np.argsort(-p[:, 4], kind="stable"):
            best, best_j = np.float32(0.0), -1
            for j in range(len(g)):
                if j in matched or int(p[idx, 5]) != int(g[j, 0]):
                    continue
                v = _iou_plain(p[idx, :4], g[j, 1:5])
                if v > best and v >= np.float32(match_iou):
                    best, best_j = v, j
            if best_j >= 0:
                matched.add(best_j)
                out.append(np.float32(1.0) - best)
    return np.asarray(out, np.float32)


def conformal_quantile(scores, alpha=0.10):
    return float(np.quantile(np.asarray(scores, np.float64), 1 - alpha))


def small_object_metric_update(preds, gts, size_threshold=15, iou_threshold=0.5, image_size=640):
    """data_loader.SmallObjectMetric.update (data_loader.py:322-389) restated with numpy fp32 scalars; returns
    (tp, fp, fn) of the batch.  preds: list of [N,6] (x_c,y_c,w,h,conf,cls), gts: list of [G,5] (cls,x_c,y_c,w,h),
    normalised.  PINNED: equal to the reference class imported from /root/reference and to
    tests/golden/small_object_metric.npz generated from it (tests/test_oracle_pins.py)."""
    f = np.float32

    def small(w, h):
        return float(w) * image_size < size_threshold and float(h) * image_size < size_threshold

    def iou(a, b):
        a1, a2, a3, a4 = a[0] - a[2] / f(2), a[1] - a[3] / f(2), a[0] + a[2] / f(2), a[1] + a[3] / f(2)
        b1, b2, b3, b4 = b[0] - b[2] / f(2), b[1] - b[3] / f(2), b[0] + b[2] / f(2), b[1] + b[3] / f(2)
        iw, ih = max(f(0), min(a3, b3) - max(a1, b1)), max(f(0), min(a4, b4) - max(a2, b2))
        inter = iw * ih
        union = (a3 - a1) * (a4 - a2) + (b3 - b1) * (b4 - b2) - inter
        return 0.0 if union <= 0 else float(inter / union)

    tp = fp = fn = 0
    for p, g in zip(preds, gts):
        p = np.asarray(p, f).reshape(-1, 6)
        g = np.asarray(g, f).reshape(-1, 5)
        sg = [r for r in g if small(r[3], r[4])]
        if not sg:
            continue
        if len(p) == 0:
            fn += len(sg)
            continue
        matched = set()
        for r in p[np.argsort(-p[:, 4], kind="stable")]:
            best, best_j = 0.0, -1
            for j, q in enumerate(sg):
                if j in matched or int(r[5]) != int(q[0]):
                    continue
                v = iou(r[:4], q[1:5])
                if v > best:
                    best, best_j = v, j
            if best >= iou_threshold:
                tp += 1
                matched.add(best_j)
            elif small(r[2], r[3]):
                fp += 1
        fn += len(sg) - len(matched)
    return tp, fp, fn
