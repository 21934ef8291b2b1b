"""Oracle post-processing: Ultralytics ``non_max_suppression`` restated (numpy + C),
and the reference's TLBR decode / greedy NMS.  TEST INFRASTRUCTURE.

Followed semantics (reference call sites train.py:396-405
(conf=0.001), eval.py:32, trainer.py:237-240 for the [N,6] column order):

  per image:  xc = amax(cls) > conf  ->  xywh->xyxy (xy -/+ wh/2)  ->  (conf, j) = max(cls)
  (first max wins)  ->  keep conf > thr  ->  if n > max_nms: stable score-descending
  top-max_nms  ->  torchvision.ops.nms(box + j*max_wh, conf, iou)  ->  [:max_det]
  ->  rows (x1,y1,x2,y2,conf,cls).

The wall-clock abort of Ultralytics (2.0 + 0.05*B s) is deliberately NOT reproduced.
Ties: ``argsort(descending=True)`` in Ultralytics is not a stable sort; the oracle freezes
"equal scores -> lower anchor index first", which is also what torchvision's nms does.
"""
from __future__ import annotations

import ctypes
from pathlib import Path

import numpy as np

from . import build as _build

_lib = None
_ref = None


class _Det(ctypes.Structure):
    _fields_ = [("x1", ctypes.c_float), ("y1", ctypes.c_float), ("x2", ctypes.c_float),
                ("y2", ctypes.c_float), ("conf", ctypes.c_float), ("cls", ctypes.c_int)]


DET_DTYPE = np.dtype([("x1", "f4"), ("y1", "f4"), ("x2", "f4"), ("y2", "f4"), ("conf", "f4"), ("cls", "i4")])


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(_build.build_oracle()))
        _lib.uydo_nms_sorted.restype = ctypes.c_int
        _lib.uydo_decode_tlbr.restype = ctypes.c_int
        _lib.uydo_nms_hpp_sorted.restype = ctypes.c_int
    return _lib


def ref_lib():
    """The reference's own postprocess.hpp, compiled (None when neither the mount nor a
    prebuilt oracle/_ref/libref_postprocess.so is available)."""
    global _ref
    if _ref is None:
        p = _build.build_ref()
        if p is None:
            return None
        _ref = ctypes.CDLL(str(p))
        _ref.ref_decode_head.restype = ctypes.c_int
        _ref.ref_nms.restype = ctypes.c_int
        _ref.ref_iou.restype = ctypes.c_float
    return _ref


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def nms_sorted(boxes_sorted: np.ndarray, iou_thr: float, max_keep: int) -> np.ndarray:
    """torchvision-semantics greedy NMS over boxes already in processing order."""
    b = np.ascontiguousarray(boxes_sorted, dtype=np.float32)
    n = b.shape[0]
    keep = np.empty(max(min(n, max_keep), 1), dtype=np.int32)
    k = lib().uydo_nms_sorted(_fp(b), ctypes.c_int(n), ctypes.c_double(iou_thr), ctypes.c_int(max_keep), _fp(keep))
    return keep[:k].copy()


def nms_torchvision_semantics(boxes: np.ndarray, scores: np.ndarray, iou_thr: float, max_keep=None) -> np.ndarray:
    """Indices (into ``boxes``) kept by torchvision.ops.nms, in score order."""
    order = np.argsort(-scores.astype(np.float32), kind="stable")
    mk = len(order) if max_keep is None else max_keep
    k = nms_sorted(boxes[order], iou_thr, mk)
    return order[k]


def non_max_suppression(pred: np.ndarray, conf_thres=0.25, iou_thres=0.45, max_det=300,
                        max_nms=30000, max_wh=7680.0, return_index=False):
    """``pred``: [B, 4+nc, A] fp32 (cx,cy,w,h, class probabilities).  Returns a list of
    [N_i, 6] fp32 arrays (and, optionally, the anchor index of every kept row)."""
    pred = np.asarray(pred, dtype=np.float32)
    B, no, A = pred.shape
    outs, idxs = [], []
    for b in range(B):
        p = pred[b].T  # [A, 4+nc]
        cls = p[:, 4:]
        xc = cls.max(1) > np.float32(conf_thres)
        anchor = np.nonzero(xc)[0]
        x = p[xc]
        half = x[:, 2:4] / np.float32(2)
        box = np.concatenate((x[:, 0:2] - half, x[:, 0:2] + half), 1).astype(np.float32)
        j = x[:, 4:].argmax(1)
        conf = x[:, 4:].max(1)
        m = conf > np.float32(conf_thres)
        box, conf, j, anchor = box[m], conf[m], j[m], anchor[m]
        order = np.argsort(-conf, kind="stable")
        if len(order) > max_nms:
            order = order[:max_nms]
        box, conf, j, anchor = box[order], conf[order], j[order], anchor[order]
        off = (j.astype(np.float32) * np.float32(max_wh))[:, None]
        keep = nms_sorted((box + off).astype(np.float32), iou_thres, max_det)
        out = np.concatenate((box[keep], conf[keep, None], j[keep, None].astype(np.float32)), 1).astype(np.float32)
        outs.append(out)
        idxs.append(anchor[keep].astype(np.int64))
    return (outs, idxs) if return_index else outs


def decode_tlbr(cls: np.ndarray, reg: np.ndarray, stride: int, conf_thr: float, q: float = 0.0,
                use_ref: bool = False, return_cells: bool = False):
    """postprocess.hpp:94-145 on one level.  cls [nc,H,W], reg [4,H,W] fp32 -> DET_DTYPE[n]."""
    cls = np.ascontiguousarray(cls, np.float32)
    reg = np.ascontiguousarray(reg, np.float32)
    nc, h, w = cls.shape
    out = np.zeros(h * w, dtype=DET_DTYPE)
    if return_cells:  # + the row-major grid cell of every detection (the restatement only)
        cells = np.zeros(h * w, dtype=np.int32)
        n = lib().uydo_decode_tlbr_cells(_fp(cls), _fp(reg), ctypes.c_int(w), ctypes.c_int(h), ctypes.c_int(stride), ctypes.c_int(nc),
                                         ctypes.c_float(conf_thr), ctypes.c_float(q), _fp(out), _fp(cells), ctypes.c_int(h * w))
        return out[:n].copy(), cells[:n].copy()
    L = ref_lib() if use_ref else lib()
    fn = L.ref_decode_head if use_ref else L.uydo_decode_tlbr
    n = fn(_fp(cls), _fp(reg), ctypes.c_int(w), ctypes.c_int(h), ctypes.c_int(stride), ctypes.c_int(nc),
           ctypes.c_float(conf_thr), ctypes.c_float(q), _fp(out), ctypes.c_int(h * w))
    return out[:n].copy()


def greedy_nms_hpp(dets: np.ndarray, iou_thr: float, use_ref: bool = False) -> np.ndarray:
    """postprocess.hpp:44-67 (class-aware greedy NMS).  Returns kept detections in order.
    The oracle sorts stably by confidence descending (the header's std::sort is unstable;
    callers of the pin tests use distinct confidences)."""
    dets = np.ascontiguousarray(dets, dtype=DET_DTYPE)
    n = len(dets)
    if use_ref:
        out = np.zeros(max(n, 1), dtype=DET_DTYPE)
        k = ref_lib().ref_nms(_fp(dets), ctypes.c_int(n), ctypes.c_float(iou_thr), _fp(out))
        return out[:k].copy()
    order = np.argsort(-dets["conf"], kind="stable")
    s = np.ascontiguousarray(dets[order])
    keep = np.empty(max(n, 1), dtype=np.int32)
    k = lib().uydo_nms_hpp_sorted(_fp(s), ctypes.c_int(n), ctypes.c_float(iou_thr), _fp(keep))
    return s[keep[:k]].copy()
