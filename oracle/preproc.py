"""Oracle: numpy restatement of the reference's camera pre-processing kernels.  TEST INFRASTRUCTURE.

Follows ros2_ws/src/perception/src/cuda_preprocess.cu: bgra_to_rgb_normalize_kernel :99-127,
resize_bgra_to_rgb_normalize_kernel :140-199, nv12_to_rgb_normalize_kernel :207-253.  Pinned on the GPU box
against the reference kernels themselves (oracle/_ref/libref_preprocess.so, built by oracle/build.py);
fp32 throughout, the compiler's multiply-add contraction is the only source of last-bit differences.
"""
from __future__ import annotations

import numpy as np

IMAGENET = (0.485, 0.456, 0.406, 0.229, 0.224, 0.225)   # NormParams() default, cuda_preprocess.cu:64-66
f32 = np.float32


def _norm(v, mean, std):
    return ((v.astype(f32) / f32(255.0)) - f32(mean)) / f32(std)


def bgra(frame: np.ndarray, params=IMAGENET) -> np.ndarray:
    """frame uint8 [H, W, 4] BGRA -> fp32 [3, H, W] RGB normalised (:99-127)."""
    mr, mg, mb, sr, sg, sb = params
    return np.stack((_norm(frame[..., 2], mr, sr), _norm(frame[..., 1], mg, sg), _norm(frame[..., 0], mb, sb)))


def bgra_resize(frame: np.ndarray, dst_h: int, dst_w: int, params=IMAGENET) -> np.ndarray:
    """Half-pixel bilinear resize with clamped source coordinates (:140-199)."""
    H, W, _ = frame.shape
    mr, mg, mb, sr, sg, sb = params
    sx = f32(W) / f32(dst_w)
    sy = f32(H) / f32(dst_h)
    x = np.clip((np.arange(dst_w, dtype=f32) + f32(0.5)) * sx - f32(0.5), f32(0), f32(W - 1)).astype(f32)
    y = np.clip((np.arange(dst_h, dtype=f32) + f32(0.5)) * sy - f32(0.5), f32(0), f32(H - 1)).astype(f32)
    x0, y0 = x.astype(np.int32), y.astype(np.int32)
    x1, y1 = np.minimum(x0 + 1, W - 1), np.minimum(y0 + 1, H - 1)
    fx, fy = (x - x0.astype(f32))[None, :], (y - y0.astype(f32))[:, None]
    w00, w01, w10, w11 = (1 - fx) * (1 - fy), fx * (1 - fy), (1 - fx) * fy, fx * fy
    out = []
    for ch, mean, std in ((2, mr, sr), (1, mg, sg), (0, mb, sb)):
        p = frame[..., ch].astype(f32)
        v = (w00 * p[y0][:, x0] + w01 * p[y0][:, x1] + w10 * p[y1][:, x0] + w11 * p[y1][:, x1]).astype(f32)
        out.append(((v / f32(255.0)) - f32(mean)) / f32(std))
    return np.stack(out).astype(f32)


def nv12(y_plane: np.ndarray, uv_plane: np.ndarray, params=IMAGENET) -> np.ndarray:
    """BT.601 NV12 -> RGB normalised (:207-253).  y [H, W], uv [H/2, W] interleaved U, V."""
    mr, mg, mb, sr, sg, sb = params
    H, W = y_plane.shape
    Y = y_plane.astype(f32)
    U = np.repeat(np.repeat(uv_plane[:, 0::2], 2, axis=0), 2, axis=1)[:H, :W].astype(f32) - f32(128)
    V = np.repeat(np.repeat(uv_plane[:, 1::2], 2, axis=0), 2, axis=1)[:H, :W].astype(f32) - f32(128)
    r = np.clip(Y + f32(1.402) * V, 0, 255).astype(f32)
    g = np.clip(Y - f32(0.344136) * U - f32(0.714136) * V, 0, 255).astype(f32)
    b = np.clip(Y + f32(1.772) * U, 0, 255).astype(f32)
    return np.stack((_norm(r, mr, sr), _norm(g, mg, sg), _norm(b, mb, sb)))
