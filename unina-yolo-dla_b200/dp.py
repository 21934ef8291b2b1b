"""Data-parallel plumbing: frames shard contiguously across ranks (one process per GPU), the
forward path has no collective; the only exchange is the gather of the fixed-shape detections
for batched evaluation (NCCL over NVLink on GPUs, gloo in the CPU tests).

The gather is ONE collective per step and never touches the host: every rank contributes the same fixed-shape
record ``[rows, max_det * 6 + 1]`` (the detections of an image followed by its count, bit-cast), shard sizes are a
function of (total, world, rank) that every rank computes locally (``shard_bounds``) instead of exchanging them,
and on CUDA the pack + ``all_gather_into_tensor`` run on a side stream so that the next step's forward overlaps
the exchange (``DetectionGather.start`` returns a handle; ``wait()`` orders the consumer after it)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [begin, end) of `total` frames owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class GatherHandle:
    """Result of DetectionGather.start: ``wait()`` makes the current stream wait for the exchange and returns
    ``(det [B_total, max_det, 6], cnt [B_total])`` in rank order."""

    def __init__(self, packed, sizes, rows, max_det, event, work=None):
        self._packed, self._sizes, self._rows, self._max_det, self._event, self._work = packed, sizes, rows, max_det, event, work

    def wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self._event is not None:
            torch.cuda.current_stream(self._packed.device).wait_event(self._event)
        p = self._packed
        if any(s != self._rows for s in self._sizes):  # ragged shards: drop the padding rows (sizes known locally)
            keep = torch.cat([torch.arange(r * self._rows, r * self._rows + s, device=p.device) for r, s in enumerate(self._sizes)])
            p = p[keep]
        det = p[:, :-1].reshape(p.shape[0], self._max_det, 6)
        cnt = p[:, -1].contiguous().view(torch.int32)
        return det, cnt


class DetectionGather:
    """Fixed-shape gather of per-image detections across the ranks of ``group``."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self._side = None
        self._bufs = {}

    def _buffers(self, key, rows, width, like):
        b = self._bufs.get(key)
        if b is None:  # two sets: step i + 1 packs while the consumer of step i may still read
            b = self._bufs[key] = ([like.new_empty(rows, width) for _ in range(2)],
                                   [like.new_empty(self.world * rows, width) for _ in range(2)], [0])
        send, recv, turn = b
        i = turn[0] = (turn[0] + 1) % 2
        return send[i], recv[i]

    def start(self, det: torch.Tensor, cnt: torch.Tensor, total: int | None = None) -> GatherHandle:
        """det [B_local, max_det, 6] fp32, cnt [B_local] int32.  ``total`` = number of frames over all ranks when
        the shards are ragged (B_local = shard_bounds(total, world, rank) size); None = equal shards."""
        max_det = det.shape[1]
        if self.world == 1:
            packed = torch.cat((det.reshape(det.shape[0], -1), cnt.view(torch.float32)[:, None]), 1)
            return GatherHandle(packed, [det.shape[0]], det.shape[0], max_det, None)
        if total is None:
            sizes = [det.shape[0]] * self.world
        else:
            sizes = [e - b for b, e in (shard_bounds(total, self.world, r) for r in range(self.world))]
            assert sizes[dist.get_rank(self.group)] == det.shape[0], "shard size does not follow shard_bounds"
        rows, width = max(sizes), max_det * 6 + 1
        send, recv = self._buffers((det.device, rows, width), rows, width, det)

        def pack():
            n = det.shape[0]
            send[:n, :-1].copy_(det.reshape(n, -1))
            send[:n, -1].copy_(cnt.view(torch.float32))
            if n < rows:
                send[n:].zero_()

        if det.is_cuda:
            if self._side is None:
                self._side = torch.cuda.Stream(device=det.device)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(det.device))
            with torch.cuda.stream(self._side):
                self._side.wait_event(ready)
                det.record_stream(self._side)
                cnt.record_stream(self._side)
                pack()
                dist.all_gather_into_tensor(recv, send, group=self.group)
                done = torch.cuda.Event()
                done.record(self._side)
            return GatherHandle(recv, sizes, rows, max_det, done)
        pack()
        work = dist.all_gather_into_tensor(recv, send, group=self.group, async_op=True)
        return GatherHandle(recv, sizes, rows, max_det, None, work)


_default = {}


def gather_detections(det: torch.Tensor, cnt: torch.Tensor, group=None, total: int | None = None):
    """Blocking convenience form: (det [B_total, max_det, 6], cnt [B_total]) on every rank, in rank order.
    No size exchange and no host synchronisation: pass ``total`` when shards are ragged."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return det, cnt
    g = _default.get(group)
    if g is None:
        g = _default[group] = DetectionGather(group)
    return g.start(det, cnt, total).wait()
