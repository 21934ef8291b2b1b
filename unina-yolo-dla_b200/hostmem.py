"""Pinned host staging buffers for frame batches (the host side of ``predict_stream``): ``pinned_frames`` returns a
torch tensor over ``cudaHostAlloc`` memory, optionally write-combined (filled sequentially by the producer, read by the
GPU over PCIe without snooping the CPU caches -- do not read it back on the CPU)."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from ._lib import check


class _HostBlock:
    def __init__(self, nbytes: int, write_combined: bool):
        self.ptr = C.c_void_p()
        check(_lib.lib().uyd_host_alloc(nbytes, int(write_combined), C.byref(self.ptr)), "uyd_host_alloc")
        self.nbytes = nbytes

    def __del__(self):
        try:
            _lib.lib().uyd_host_free(self.ptr)
        except Exception:
            pass


def pinned_frames(shape, dtype=torch.uint8, write_combined: bool = True) -> torch.Tensor:
    nbytes = math.prod(shape) * torch.empty((), dtype=dtype).element_size()
    blk = _HostBlock(nbytes, write_combined)
    buf = (C.c_uint8 * nbytes).from_address(blk.ptr.value)
    t = torch.frombuffer(buf, dtype=dtype).view(*shape)
    t._uyd_block = blk   # keeps the allocation alive as long as the tensor
    return t
