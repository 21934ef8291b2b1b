"""Thin object layer over the libuyd plan API (include/uyd.h): buffers, channel slices, ops."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import (CHAIN_DFL, CHAIN_PW3, CHAIN_STORE, IMPL_AUTO, UYD_BF16, UYD_F32, UYD_S8, UYD_U8, C3kDesc, ChainDesc,
                   ClsBranchDesc, ConvDesc, ConvS8Desc, check)

_TORCH_DTYPE = {UYD_BF16: torch.bfloat16, UYD_F32: torch.float32, UYD_S8: torch.int8}


@dataclass(frozen=True)
class Slice:
    """Channels [coff, coff + c) of activation buffer `buf` ([B, h, w, C_total] NHWC)."""
    buf: int
    coff: int
    c: int
    h: int
    w: int

    def sub(self, off: int, c: int) -> "Slice":
        assert 0 <= off and off + c <= self.c
        return Slice(self.buf, self.coff + off, c, self.h, self.w)


NETWORK_INPUT = Slice(-1, 0, 3, 0, 0)


def fold_bn(conv: torch.nn.Conv2d, bn: torch.nn.BatchNorm2d):
    """fuse_conv_and_bn: W' = diag(g/sqrt(var+eps)) W, b' = beta - mu*g/sqrt(var+eps)."""
    w = conv.weight.detach().double().cpu()
    s = bn.weight.detach().double().cpu() / torch.sqrt(bn.running_var.detach().double().cpu() + bn.eps)
    b = bn.bias.detach().double().cpu() - bn.running_mean.detach().double().cpu() * s
    if conv.bias is not None:
        b = b + conv.bias.detach().double().cpu() * s
    return (w * s.view(-1, 1, 1, 1)).float().numpy(), b.float().numpy()


class Plan:
    def __init__(self, device: int, max_batch: int):
        self.device = device
        self.max_batch = max_batch
        self.ctx = _lib.context(device)
        self.handle = C.c_void_p()
        check(_lib.lib().uyd_plan_create(self.ctx, max_batch, C.byref(self.handle)), "uyd_plan_create")
        self.shapes = {}
        self.heads = []
        self.finalized = False

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().uyd_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- construction ------------------------------------------------------------------
    def buffer(self, h: int, w: int, c: int, dtype: int = UYD_BF16) -> Slice:
        bid = C.c_int()
        check(_lib.lib().uyd_plan_add_buffer(self.handle, h, w, c, dtype, C.byref(bid)), "uyd_plan_add_buffer")
        self.shapes[bid.value] = (h, w, c, dtype)
        return Slice(bid.value, 0, c, h, w)

    def buffer_slice(self, buf: int, coff: int, c: int) -> Slice:
        h, w, ctot, _ = self.shapes[buf]
        assert coff + c <= ctot
        return Slice(buf, coff, c, h, w)

    def conv(self, src: Slice, dst: Slice, weight: np.ndarray, bias: np.ndarray, k: int, stride: int = 1,
             relu: bool = True, depthwise: bool = False, res: Slice | None = None, impl: int = IMPL_AUTO,
             pre: Slice | None = None) -> Slice:
        """``pre``: an fp32 buffer [h/2, w/2, cout] of partial sums that is added, nearest-x2 upsampled, before the
        activation (Upsample + Concat + 1x1 conv without materialising the upsampled tensor)."""
        weight = np.ascontiguousarray(weight, dtype=np.float32)
        bias = np.ascontiguousarray(bias, dtype=np.float32)
        cout = weight.shape[0]
        cin = cout if depthwise else weight.shape[1]
        assert weight.shape[2] == k and weight.shape[3] == k and bias.shape == (cout,)
        assert dst.c == cout and (src.buf < 0 or src.c == cin), (src, dst, weight.shape)
        d = ConvDesc(src.buf, src.coff, dst.buf, dst.coff, res.buf if res else -1, res.coff if res else 0,
                     cin, cout, k, stride, int(depthwise), int(relu), impl, (pre.buf + 1) if pre is not None else 0)
        check(_lib.lib().uyd_plan_add_conv(self.handle, C.byref(d), weight.ctypes.data_as(C.c_void_p),
                                           bias.ctypes.data_as(C.c_void_p)), "uyd_plan_add_conv")
        return dst

    def conv_s8(self, src: Slice, dst: Slice, weight_q: np.ndarray, mult: np.ndarray, bias: np.ndarray, k: int,
                stride: int = 1, relu: bool = True, out_scale: float = 0.0, impl: int = IMPL_AUTO, depthwise: bool = False,
                res: Slice | None = None, out_round_bf16: bool = False, pre: Slice | None = None) -> Slice:
        """INT8 conv: src in a UYD_S8 buffer, weight_q int8 [cout][cin][k][k] (depth-wise: [c][1][k][k]); the dtype
        of dst's buffer selects the epilogue (int8 re-quantised with out_scale, or fp32 / bf16); ``res`` is a bf16
        slice added after the activation."""
        weight_q = np.ascontiguousarray(weight_q, dtype=np.int8)
        mult = np.ascontiguousarray(mult, dtype=np.float32)
        bias = np.ascontiguousarray(bias, dtype=np.float32)
        cout = weight_q.shape[0]
        cin = cout if depthwise else weight_q.shape[1]
        assert dst.c == cout and src.c == cin and mult.shape == (cout,) and bias.shape == (cout,)
        d = ConvS8Desc(src.buf, src.coff, dst.buf, dst.coff, cin, cout, k, stride, int(relu), float(out_scale), impl,
                       int(depthwise), res.buf if res else -1, res.coff if res else 0, int(out_round_bf16),
                       (pre.buf + 1) if pre is not None else 0)
        check(_lib.lib().uyd_plan_add_conv_s8(self.handle, C.byref(d), weight_q.ctypes.data_as(C.c_void_p),
                                              mult.ctypes.data_as(C.c_void_p), bias.ctypes.data_as(C.c_void_p)),
              "uyd_plan_add_conv_s8")
        return dst

    def quantize(self, src: Slice, dst: Slice, scale: float) -> Slice:
        """bf16 slice -> int8 slice, q = clamp(rne(x * scale), -127, 127) (uyd_plan_add_quantize)."""
        assert src.c == dst.c
        check(_lib.lib().uyd_plan_add_quantize(self.handle, src.buf, src.coff, dst.buf, dst.coff, src.c, float(scale)),
              "uyd_plan_add_quantize")
        return dst

    def slice_absmax(self, s: Slice, batch: int, out_bits: torch.Tensor) -> None:
        """atomicMax of the float bits of max|x| over a bf16 slice into ``out_bits`` (uint32/int32 device scalar)."""
        check(_lib.lib().uyd_plan_slice_absmax(self.handle, s.buf, s.coff, s.c, batch, C.c_void_p(out_bits.data_ptr()),
                                               self._stream()), "uyd_plan_slice_absmax")

    def slice_histogram(self, s: Slice, batch: int, inv_width: float, hist: torch.Tensor) -> None:
        """Adds the histogram of |x| over a bf16 slice into ``hist`` (int32 device vector, one entry per bin)."""
        check(_lib.lib().uyd_plan_slice_histogram(self.handle, s.buf, s.coff, s.c, batch, inv_width, hist.numel(),
                                                  C.c_void_p(hist.data_ptr()), self._stream()), "uyd_plan_slice_histogram")

    C3K_WIDTHS = (8, 16, 32)

    @staticmethod
    def c3k_supported(src: Slice, dst: Slice, src_pitch_ok: bool = True) -> bool:
        """Shape rule of the fused C3k kernel (mirrors c3k_supported in csrc/c3k_fused.cu)."""
        th_ok = src.h % 32 == 0 or src.h % 20 == 0 or src.h % 16 == 0
        return src.c in Plan.C3K_WIDTHS and src.w % 40 == 0 and th_ok and src.coff % 8 == 0 and dst.coff % 8 == 0

    def c3k(self, src: Slice, dst: Slice, weights: list, biases: list) -> Slice:
        """Fused C3k block: weights/biases = [cv1, cv2, m0.cv1, m0.cv2, m1.cv1, m1.cv2, cv3], BN folded."""
        ws = [np.ascontiguousarray(w, dtype=np.float32) for w in weights]
        bs = [np.ascontiguousarray(b, dtype=np.float32) for b in biases]
        assert len(ws) == 7 and len(bs) == 7
        wp = (C.c_void_p * 7)(*[w.ctypes.data_as(C.c_void_p) for w in ws])
        bp = (C.c_void_p * 7)(*[b.ctypes.data_as(C.c_void_p) for b in bs])
        d = C3kDesc(src.buf, src.coff, dst.buf, dst.coff, src.c, 0)
        check(_lib.lib().uyd_plan_add_c3k(self.handle, C.byref(d), wp, bp), "uyd_plan_add_c3k")
        return dst

    def c3k_s8(self, src: Slice, dst: Slice, weights_q: list, mults: list, biases: list, in_scales: list) -> Slice:
        """Fused C3k block of the INT8 graph (uyd_plan_add_c3k_s8): per conv [cv1, cv2, m0.cv1, m0.cv2, m1.cv1, m1.cv2, cv3]
        its int8 weight codes, requant pair (m_c, b_c) and input scale; src / dst are bf16 slices."""
        ws = [np.ascontiguousarray(w, dtype=np.int8) for w in weights_q]
        ms = [np.ascontiguousarray(m, dtype=np.float32) for m in mults]
        bs = [np.ascontiguousarray(b, dtype=np.float32) for b in biases]
        assert len(ws) == 7 and len(ms) == 7 and len(bs) == 7 and len(in_scales) == 7
        wp = (C.c_void_p * 7)(*[w.ctypes.data_as(C.c_void_p) for w in ws])
        mp = (C.c_void_p * 7)(*[m.ctypes.data_as(C.c_void_p) for m in ms])
        bp = (C.c_void_p * 7)(*[b.ctypes.data_as(C.c_void_p) for b in bs])
        sc = (C.c_float * 7)(*[float(x) for x in in_scales])
        d = C3kDesc(src.buf, src.coff, dst.buf, dst.coff, src.c, 0)
        check(_lib.lib().uyd_plan_add_c3k_s8(self.handle, C.byref(d), wp, mp, bp, sc), "uyd_plan_add_c3k_s8")
        return dst

    @staticmethod
    def cls_branch_supported(src: Slice, mid: int, nc: int) -> bool:
        return src.c in (32, 64) and mid == 32 and nc <= 8 and src.w % 40 == 0 and src.h % 8 == 0 and src.coff % 8 == 0

    def cls_branch(self, src: Slice, dst: Slice, mid: int, weights: list, biases: list) -> Slice:
        """Fused Detect class branch: weights/biases = [dw1, pw1, dw2, pw2, pw3] (BN folded); dst = fp32 head slice."""
        ws = [np.ascontiguousarray(w, dtype=np.float32) for w in weights]
        bs = [np.ascontiguousarray(b, dtype=np.float32) for b in biases]
        assert len(ws) == 5 and len(bs) == 5
        wp = (C.c_void_p * 5)(*[w.ctypes.data_as(C.c_void_p) for w in ws])
        bp = (C.c_void_p * 5)(*[b.ctypes.data_as(C.c_void_p) for b in bs])
        d = ClsBranchDesc(src.buf, src.coff, dst.buf, dst.coff, src.c, mid, dst.c, 0)
        check(_lib.lib().uyd_plan_add_cls_branch(self.handle, C.byref(d), wp, bp), "uyd_plan_add_cls_branch")
        return dst

    def stem2(self, dst: Slice, w0, b0, w1, b1, w2=None, b2=None) -> Slice:
        """Fused Conv(3,16,3,2) -> Conv(16,32,3,2) reading the network input (uyd_plan_add_stem2); with
        ``w2, b2`` the 1x1 Conv(32,16) consuming it runs in the same launch (uyd_plan_add_stem2_pw)."""
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        w0, b0, w1, b1 = f32(w0), f32(b0), f32(w1), f32(b1)
        assert w0.shape == (16, 3, 3, 3) and w1.shape == (32, 16, 3, 3)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        if w2 is None:
            assert dst.c == 32
            check(_lib.lib().uyd_plan_add_stem2(self.handle, dst.buf, dst.coff, ptr(w0), ptr(b0), ptr(w1), ptr(b1)), "uyd_plan_add_stem2")
        else:
            w2, b2 = f32(w2).reshape(16, 32), f32(b2)
            assert dst.c == 16 and b2.shape == (16,)
            check(_lib.lib().uyd_plan_add_stem2_pw(self.handle, dst.buf, dst.coff, ptr(w0), ptr(b0), ptr(w1), ptr(b1), ptr(w2), ptr(b2)),
                  "uyd_plan_add_stem2_pw")
        return dst

    @staticmethod
    def chain_supported(src: Slice, n1: int, n2: int) -> bool:
        """Shape rule of the chained head kernel (chain_supported in csrc/conv_chain.cu)."""
        return src.c in (32, 64) and n1 in (32, 64) and 1 <= n2 <= 64 and src.coff % 8 == 0

    def chain(self, src: Slice, w1, b1, w2, b2, *, dw1: bool = False, relu2: bool = False, final: int = CHAIN_STORE,
              out: Slice | None = None, w3=None, b3=None, a_total: int = 0, a_off: int = 0, y_ch0: int = 0, no: int = 0,
              stride: float = 1.0) -> Slice | None:
        """3x3 Conv+BN+ReLU (dense or depth-wise) -> 1x1 conv -> fused final stage, one launch
        (uyd_plan_add_chain).  Weights BN-folded, PyTorch layout."""
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        w1, b1, w2, b2 = f32(w1), f32(b1), f32(w2).reshape(np.shape(w2)[0], -1), f32(b2)
        n1, n2 = w1.shape[0], w2.shape[0]
        assert w2.shape[1] == n1 and b1.shape == (n1,) and b2.shape == (n2,)
        assert w1.shape[1] == (1 if dw1 else src.c) and w1.shape[2:] == (3, 3)
        nc = 0
        if final == CHAIN_PW3:
            w3, b3 = f32(w3).reshape(np.shape(w3)[0], -1), f32(b3)
            nc = w3.shape[0]
            assert w3.shape[1] == n2 and b3.shape == (nc,)
        d = ChainDesc(src.buf, src.coff, src.c, n1, n2, int(dw1), int(relu2), final, out.buf if out else -1,
                      out.coff if out else 0, nc, a_total, a_off, y_ch0, no, float(stride))
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        check(_lib.lib().uyd_plan_add_chain(self.handle, C.byref(d), ptr(w1), ptr(b1), ptr(w2), ptr(b2),
                                            ptr(w3) if final == CHAIN_PW3 else None, ptr(b3) if final == CHAIN_PW3 else None),
              "uyd_plan_add_chain")
        return out

    def sppf_pool(self, s: Slice, c: int) -> None:
        check(_lib.lib().uyd_plan_add_sppf_pool(self.handle, s.buf, s.coff, c), "uyd_plan_add_sppf_pool")

    def upsample2x(self, src: Slice, dst: Slice) -> Slice:
        check(_lib.lib().uyd_plan_add_upsample2x(self.handle, src.buf, src.coff, dst.buf, dst.coff, src.c),
              "uyd_plan_add_upsample2x")
        return dst

    def set_heads(self, heads: list[Slice], strides: list[int], reg_max: int, nc: int) -> None:
        n = len(heads)
        ids = (C.c_int * n)(*[h.buf for h in heads])
        st = (C.c_int * n)(*strides)
        check(_lib.lib().uyd_plan_set_heads(self.handle, ids, st, n, reg_max, nc), "uyd_plan_set_heads")
        self.heads = heads

    def finalize(self) -> "Plan":
        check(_lib.lib().uyd_plan_finalize(self.handle), "uyd_plan_finalize")
        self.finalized = True
        return self

    # ---- execution ---------------------------------------------------------------------
    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)   # the plan's device, not the current one

    def run(self, x: torch.Tensor, y: torch.Tensor | None = None) -> None:
        """Runs every op; ``y`` ([B, no, A] fp32) receives the decoded prediction of plans whose head
        kernels decode in their epilogue."""
        assert x.is_cuda and x.is_contiguous() and x.dtype in (torch.float32, torch.uint8)
        if y is not None:
            assert y.is_cuda and y.is_contiguous() and y.dtype == torch.float32 and y.shape[0] >= x.shape[0]
            check(_lib.lib().uyd_plan_run_decoded(self.handle, C.c_void_p(x.data_ptr()),
                                                  UYD_F32 if x.dtype == torch.float32 else UYD_U8, x.shape[0],
                                                  C.c_void_p(y.data_ptr()), self._stream()), "uyd_plan_run_decoded")
            return
        fn = _lib.lib().uyd_plan_run if x.dtype == torch.float32 else _lib.lib().uyd_plan_run_u8
        check(fn(self.handle, C.c_void_p(x.data_ptr()), x.shape[0], self._stream()), "uyd_plan_run")

    def run_camera(self, frames: "_lib.CameraFrames", batch: int, y: torch.Tensor | None = None) -> None:
        """Runs every op on camera frames (packed BGRA / NV12 device bytes): the stem normalises and resamples on load."""
        check(_lib.lib().uyd_plan_run_camera(self.handle, C.byref(frames), batch, C.c_void_p(y.data_ptr()) if y is not None else None,
                                             self._stream()), "uyd_plan_run_camera")

    def profile(self, x: torch.Tensor, y: torch.Tensor | None = None) -> list[float]:
        """Per-op milliseconds of one pass (CUDA events around every op)."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        check(_lib.lib().uyd_plan_set_profile_output(self.handle, C.c_void_p(y.data_ptr()) if y is not None else None),
              "uyd_plan_set_profile_output")
        ms = (C.c_float * self.launches)()
        check(_lib.lib().uyd_plan_profile(self.handle, C.c_void_p(x.data_ptr()), x.shape[0], self._stream(), ms),
              "uyd_plan_profile")
        return list(ms)

    def op_info(self, op: int):
        buf = C.create_string_buffer(160)
        fl, by = C.c_double(), C.c_double()
        check(_lib.lib().uyd_plan_op_info(self.handle, op, buf, 160, C.byref(fl), C.byref(by)), "uyd_plan_op_info")
        return buf.value.decode(), fl.value, by.value

    def set_timed_op(self, op: int, max_samples: int = 256) -> None:
        check(_lib.lib().uyd_plan_set_timed_op(self.handle, op, max_samples), "uyd_plan_set_timed_op")

    def timed_op_read(self):
        tot, n = C.c_float(), C.c_int()
        check(_lib.lib().uyd_plan_timed_op_read(self.handle, C.byref(tot), C.byref(n)), "uyd_plan_timed_op_read")
        return tot.value, n.value

    def run_no_input(self, batch: int, y: torch.Tensor | None = None) -> None:
        """Runs a plan whose first op does not read the network input (layer-level tests)."""
        if y is not None:
            dummy = torch.empty(16, dtype=torch.float32, device=y.device)
            check(_lib.lib().uyd_plan_run_decoded(self.handle, C.c_void_p(dummy.data_ptr()), UYD_F32, batch,
                                                  C.c_void_p(y.data_ptr()), self._stream()), "uyd_plan_run_decoded")
            return
        check(_lib.lib().uyd_plan_run(self.handle, None, batch, self._stream()), "uyd_plan_run")

    def write(self, s: Slice, nchw: torch.Tensor) -> None:
        """Stores an NCHW tensor into a slice of a bf16 buffer (layer-level tests)."""
        h, w, ctot, dtype = self.shapes[s.buf]
        tdt = _TORCH_DTYPE[dtype]
        batch = nchw.shape[0]
        ptr = C.c_void_p()
        check(_lib.lib().uyd_plan_buffer_ptr(self.handle, s.buf, C.byref(ptr)), "uyd_plan_buffer_ptr")
        full = torch.empty(batch, h, w, ctot, dtype=tdt, device=f"cuda:{self.device}")
        nbytes = full.numel() * full.element_size()
        check(_lib.lib().uyd_memcpy_d2d(C.c_void_p(full.data_ptr()), ptr, nbytes, self._stream()), "uyd_memcpy_d2d")
        full[..., s.coff:s.coff + s.c] = nchw.to(full.device).permute(0, 2, 3, 1).to(tdt)
        check(_lib.lib().uyd_memcpy_d2d(ptr, C.c_void_p(full.data_ptr()), nbytes, self._stream()), "uyd_memcpy_d2d")
        torch.cuda.current_stream(self.device).synchronize()

    def decode(self, y: torch.Tensor, batch: int) -> None:
        check(_lib.lib().uyd_plan_run_decode(self.handle, C.c_void_p(y.data_ptr()), batch, self._stream()),
              "uyd_plan_run_decode")

    def export_head(self, level: int, out: torch.Tensor, batch: int) -> None:
        check(_lib.lib().uyd_plan_export_head_nchw(self.handle, level, C.c_void_p(out.data_ptr()), batch, self._stream()),
              "uyd_plan_export_head_nchw")

    def read(self, s: Slice, batch: int) -> torch.Tensor:
        """Copies a slice out as an NCHW fp32 tensor (layer-level parity checks)."""
        h, w, ctot, dtype = self.shapes[s.buf]
        tdt = _TORCH_DTYPE[dtype]
        full = torch.empty(batch, h, w, ctot, dtype=tdt, device=f"cuda:{self.device}")
        ptr = C.c_void_p()
        check(_lib.lib().uyd_plan_buffer_ptr(self.handle, s.buf, C.byref(ptr)), "uyd_plan_buffer_ptr")
        check(_lib.lib().uyd_memcpy_d2d(C.c_void_p(full.data_ptr()), ptr, full.numel() * full.element_size(), self._stream()),
              "uyd_memcpy_d2d")
        out = full[..., s.coff:s.coff + s.c].permute(0, 3, 1, 2).contiguous()
        return out if dtype == UYD_S8 else out.float()

    @property
    def bytes(self) -> int:
        return int(_lib.lib().uyd_plan_bytes(self.handle))

    @property
    def launches(self) -> int:
        return int(_lib.lib().uyd_plan_num_launches(self.handle))
