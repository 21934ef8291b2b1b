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
                                         C.c_void_p(scores.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
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
