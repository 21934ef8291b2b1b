"""Data-parallel plumbing: frames shard contiguously across ranks (one process per GPU), the
forward path has no collective; the only exchange is the gather of the fixed-shape detections
for batched evaluation (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [begin, end) of `total` frames owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_detections(det: torch.Tensor, cnt: torch.Tensor, group=None):
    """det [B_local, max_det, 6], cnt [B_local] -> (det [B_total, max_det, 6], cnt [B_total]) on every
    rank, in rank order.  Shards may differ in size by one frame (padded for the collective)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return det, cnt
    world = dist.get_world_size(group)
    n = torch.tensor([det.shape[0]], dtype=torch.int64, device=det.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    if det.shape[0] < m:
        det = torch.cat((det, det.new_zeros(m - det.shape[0], *det.shape[1:])))
        cnt = torch.cat((cnt, cnt.new_zeros(m - cnt.shape[0])))
    gd = det.new_empty(world * m, *det.shape[1:])
    gc = cnt.new_empty(world * m)
    dist.all_gather_into_tensor(gd, det.contiguous(), group=group)
    dist.all_gather_into_tensor(gc, cnt.contiguous(), group=group)
    keep = torch.cat([torch.arange(r * m, r * m + s, device=det.device) for r, s in enumerate(sizes)])
    return gd[keep], gc[keep]
