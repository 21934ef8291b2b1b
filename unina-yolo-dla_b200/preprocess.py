"""Camera-frame pre-processing on the GPU (reference: ros2_ws/src/perception/src/cuda_preprocess.cu:99-323,
the step in front of the hot path): packed BGRA / NV12 bytes -> the NCHW fp32 tensor ``UninaYoloB200`` consumes."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import NormParams, check


def norm_params(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)) -> NormParams:
    """create_norm_params (cuda_preprocess.h:55-56); the defaults are ImageNet (create_norm_params_imagenet)."""
    return NormParams(*mean, *std)


UNIT = NormParams(0.0, 0.0, 0.0, 1.0, 1.0, 1.0)   # plain x / 255


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def bgra(frames: torch.Tensor, params: NormParams = UNIT, size: tuple[int, int] | None = None) -> torch.Tensor:
    """``frames``: uint8 ``[B, H, W, 4]`` (BGRA, device) -> fp32 ``[B, 3, H', W']`` RGB, ``(x/255 - mean)/std``;
    ``size=(H', W')`` resizes with half-pixel bilinear sampling (preprocess_bgra_resize), else H' = H, W' = W."""
    assert frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 4 and frames.is_contiguous()
    B, H, W, _ = frames.shape
    oh, ow = size or (H, W)
    out = torch.empty(B, 3, oh, ow, dtype=torch.float32, device=frames.device)
    L = _lib.lib()
    with torch.cuda.device(frames.device):   # the reference entry points carry no device: they run on the current one
        return _bgra(L, frames, out, params, size, B, H, W, oh, ow)


def _bgra(L, frames, out, params, size, B, H, W, oh, ow):
    if size is None or (oh, ow) == (H, W):
        check(L.uyd_preprocess_bgra_batch(C.c_void_p(frames.data_ptr()), C.c_void_p(out.data_ptr()), B, H * W * 4, W, H, W * 4,
                                          params, _stream(frames)), "uyd_preprocess_bgra_batch")
    else:
        check(L.uyd_preprocess_bgra_resize_batch(C.c_void_p(frames.data_ptr()), C.c_void_p(out.data_ptr()), B, H * W * 4, W, H,
                                                 W * 4, ow, oh, params, _stream(frames)), "uyd_preprocess_bgra_resize_batch")
    return out


def nv12(y_plane: torch.Tensor, uv_plane: torch.Tensor, params: NormParams = UNIT) -> torch.Tensor:
    """One NV12 frame: ``y_plane`` uint8 ``[H, W]``, ``uv_plane`` uint8 ``[H/2, W]`` (interleaved U, V) -> ``[1, 3, H, W]``."""
    assert y_plane.is_cuda and uv_plane.is_cuda and y_plane.dtype == torch.uint8 and uv_plane.dtype == torch.uint8
    H, W = y_plane.shape
    out = torch.empty(1, 3, H, W, dtype=torch.float32, device=y_plane.device)
    with torch.cuda.device(y_plane.device):
        check(_lib.lib().uyd_preprocess_nv12(C.c_void_p(y_plane.data_ptr()), C.c_void_p(uv_plane.data_ptr()), C.c_void_p(out.data_ptr()),
                                             W, H, y_plane.stride(0), uv_plane.stride(0), params, _stream(y_plane)), "uyd_preprocess_nv12")
    return out
