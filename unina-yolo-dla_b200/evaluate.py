"""Batched evaluation consumer of the detections (reference: UninaValidator trainer.py:196-286, the
small-object metrics; calibrate_conformal_prediction train.py:299-520, the conformal quantile).  Everything
stays on the GPU; with several ranks the detections are gathered first (dp.gather_detections, the only
collective of the path) or the counters / scores are reduced at the end (``reduce``)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check


class DetectionEvaluator:
    def __init__(self, size_threshold: float = 15.0, small_iou: float = 0.45, match_iou: float = 0.5, device=None):
        self.size_threshold, self.small_iou, self.match_iou = float(size_threshold), float(small_iou), float(match_iou)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.reset()

    def reset(self) -> None:
        self.counters = torch.zeros(3, dtype=torch.int64, device=self.device)   # small-object TP, FP, FN
        self._scores = []

    @torch.no_grad()
    def update(self, det: torch.Tensor, cnt: torch.Tensor, gt: torch.Tensor, gt_cnt: torch.Tensor) -> None:
        """det [B, max_det, 6] / cnt [B] as returned by ``predict_batched``; gt [B, G, 5] rows
        (cls, x1, y1, x2, y2) in pixels, gt_cnt [B] (int32).  No host synchronisation."""
        assert det.is_cuda and det.dtype == torch.float32 and det.is_contiguous() and det.shape[2] == 6
        assert gt.is_cuda and gt.dtype == torch.float32 and gt.is_contiguous() and gt.shape[2] == 5
        assert cnt.dtype == torch.int32 and gt_cnt.dtype == torch.int32 and cnt.is_cuda and gt_cnt.is_cuda
        B, max_det, _ = det.shape
        scores = torch.empty(B, max_det, dtype=torch.float32, device=det.device)
        dev = det.device.index if det.device.index is not None else torch.cuda.current_device()
        check(_lib.lib().uyd_eval_update(_lib.context(dev), C.c_void_p(det.data_ptr()), C.c_void_p(cnt.data_ptr()), B, max_det,
                                         C.c_void_p(gt.data_ptr()), C.c_void_p(gt_cnt.data_ptr()), gt.shape[1], self.size_threshold,
                                         self.small_iou, self.match_iou, C.c_void_p(self.counters.data_ptr()),
                                         C.c_void_p(scores.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
              "uyd_eval_update")
        self._scores.append(scores)

    def reduce(self, group=None) -> None:
        """Sums the counters and concatenates the scores over the ranks (when every rank evaluated its own shard)."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        dist.all_reduce(self.counters, group=group)
        s = torch.cat([t.flatten() for t in self._scores]) if self._scores else torch.empty(0, device=self.device)
        n = torch.tensor([s.numel()], device=self.device)
        sizes = [torch.zeros_like(n) for _ in range(dist.get_world_size(group))]
        dist.all_gather(sizes, n, group=group)
        m = int(max(int(x) for x in sizes))
        pad = torch.full((m,), -1.0, device=self.device)
        pad[: s.numel()] = s
        out = [torch.empty_like(pad) for _ in sizes]
        dist.all_gather(out, pad, group=group)
        self._scores = [torch.cat(out)]

    def small_object_metrics(self) -> dict:
        """trainer.py:267-285 (the 1e-7 guards included)."""
        tp, fp, fn = (int(v) for v in self.counters.tolist())
        precision = tp / (tp + fp + 1e-7)
        recall = tp / (tp + fn + 1e-7)
        f1 = 2 * (precision * recall) / (precision + recall + 1e-7)
        return {"metrics/small_precision": precision, "metrics/small_recall": recall, "metrics/small_f1": f1,
                "tp": tp, "fp": fp, "fn": fn}

    def nonconformity_scores(self) -> torch.Tensor:
        if not self._scores:
            return torch.empty(0, device=self.device)
        s = torch.cat([t.flatten() for t in self._scores])
        return s[s >= 0]

    def conformal(self, alpha: float = 0.10) -> dict:
        """train.py:491-512: q_hat = the (1 - alpha) quantile (linear interpolation, numpy's default) of 1 - IoU."""
        s = self.nonconformity_scores().double()
        if s.numel() == 0:
            raise ValueError("Conformal Prediction Calibration failed: No matched predictions found.")
        srt, _ = torch.sort(s)
        pos = (srt.numel() - 1) * (1 - alpha)
        lo = int(pos // 1)
        hi = min(lo + 1, srt.numel() - 1)
        q_hat = float(srt[lo] + (srt[hi] - srt[lo]) * (pos - lo))
        return {"alpha": alpha, "coverage_target": 1 - alpha, "q_hat": q_hat, "dilation_factor": q_hat,
                "num_calibration_samples": int(srt.numel()), "mean_nonconformity": float(s.mean()),
                "std_nonconformity": float(s.std(unbiased=False))}


class SmallObjectMetric:
    """Drop-in for ``data_loader.SmallObjectMetric`` (data_loader.py:249-414): same constructor, ``reset`` /
    ``update(predictions, ground_truths)`` / ``compute``, same keys; the matching runs on the GPU, one CTA per
    image (csrc/evalmatch.cu ``small_object_metric_kernel``).  ``update`` takes the reference's lists
    (``[N,6]`` rows x_c,y_c,w,h,conf,cls and ``[G,5]`` rows cls,x_c,y_c,w,h, normalised) or already padded device
    tensors through ``update_batched``; ``from_xyxy_pixels`` converts ``predict`` rows."""

    def __init__(self, size_threshold: int = 15, iou_threshold: float = 0.5, image_size: int = 640, device=None) -> None:
        self.size_threshold, self.iou_threshold, self.image_size = size_threshold, iou_threshold, image_size
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.reset()

    def reset(self) -> None:
        self._counters = torch.zeros(3, dtype=torch.int64, device=self.device)

    @property
    def true_positives(self) -> int:
        return int(self._counters[0])

    @property
    def false_positives(self) -> int:
        return int(self._counters[1])

    @property
    def false_negatives(self) -> int:
        return int(self._counters[2])

    @staticmethod
    def from_xyxy_pixels(det: torch.Tensor, image_size: float) -> torch.Tensor:
        """[..., 6] rows (x1,y1,x2,y2,conf,cls) in pixels -> (x_c,y_c,w,h,conf,cls) normalised."""
        out = det.clone()
        out[..., 0] = (det[..., 0] + det[..., 2]) / 2 / image_size
        out[..., 1] = (det[..., 1] + det[..., 3]) / 2 / image_size
        out[..., 2] = (det[..., 2] - det[..., 0]) / image_size
        out[..., 3] = (det[..., 3] - det[..., 1]) / image_size
        return out

    @torch.no_grad()
    def update_batched(self, pred: torch.Tensor, cnt: torch.Tensor, gt: torch.Tensor, gt_cnt: torch.Tensor) -> None:
        """pred [B, max_det, 6] in confidence order, cnt [B] int32, gt [B, G, 5], gt_cnt [B] int32: device tensors."""
        assert pred.is_cuda and pred.dtype == torch.float32 and pred.is_contiguous() and pred.shape[2] == 6
        assert gt.is_cuda and gt.dtype == torch.float32 and gt.is_contiguous() and gt.shape[2] == 5
        assert cnt.dtype == torch.int32 and gt_cnt.dtype == torch.int32
        dev = pred.device.index if pred.device.index is not None else torch.cuda.current_device()
        check(_lib.lib().uyd_small_object_metric_update(
            _lib.context(dev), C.c_void_p(pred.data_ptr()), C.c_void_p(cnt.data_ptr()), pred.shape[0], pred.shape[1],
            C.c_void_p(gt.data_ptr()), C.c_void_p(gt_cnt.data_ptr()), gt.shape[1], float(self.size_threshold), float(self.iou_threshold),
            float(self.image_size), C.c_void_p(self._counters.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
            "uyd_small_object_metric_update")

    @torch.no_grad()
    def update(self, predictions, ground_truths) -> None:
        B = len(predictions)
        if B == 0:
            return
        md = max(1, max(int(p.shape[0]) if p.numel() else 0 for p in predictions))
        gm = max(1, max(int(g.shape[0]) if g.numel() else 0 for g in ground_truths))
        pred = torch.zeros(B, md, 6, device=self.device)
        gt = torch.zeros(B, gm, 5, device=self.device)
        cnt = torch.zeros(B, dtype=torch.int32, device=self.device)
        gcnt = torch.zeros(B, dtype=torch.int32, device=self.device)
        for i, (p, g) in enumerate(zip(predictions, ground_truths)):
            if p.numel():
                p = p.to(self.device, torch.float32)
                order = torch.sort(p[:, 4], descending=True, stable=True).indices   # data_loader.py:349-350
                pred[i, : p.shape[0]] = p[order]
                cnt[i] = p.shape[0]
            if g.numel():
                gt[i, : g.shape[0]] = g.to(self.device, torch.float32)
                gcnt[i] = g.shape[0]
        self.update_batched(pred, cnt, gt, gcnt)

    def compute(self) -> dict:
        tp, fp, fn = (int(v) for v in self._counters.tolist())
        precision = tp / (tp + fp) if (tp + fp) > 0 else 0.0
        recall = tp / (tp + fn) if (tp + fn) > 0 else 0.0
        f1 = 2 * (precision * recall) / (precision + recall) if (precision + recall) > 0 else 0.0
        return {"small_object_precision": precision, "small_object_recall": recall, "small_object_f1": f1,
                "small_object_tp": tp, "small_object_fp": fp, "small_object_fn": fn}
