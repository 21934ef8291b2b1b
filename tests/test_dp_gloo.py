"""N > 1 path on CPU: world_size-2 gloo processes shard a batch, post-process their shard and
gather the detections; the result must equal the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unina_yolo_dla_b200.dp import gather_detections, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _pack(rows, max_det=300):
    det = torch.zeros(len(rows), max_det, 6)
    cnt = torch.zeros(len(rows), dtype=torch.int32)
    for i, r in enumerate(rows):
        det[i, : len(r)] = torch.from_numpy(r)
        cnt[i] = len(r)
    return det, cnt


def _worker(rank, world, port, total, out_path):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from oracle import postproc as pp
    import uyd_testlib_cpu as T

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    y = T.synth_predictions(total, 4, 2100, seed=3, frac_conf=0.3)
    b, e = shard_bounds(total, world, rank)
    det, cnt = _pack(pp.non_max_suppression(y[b:e], 0.25, 0.7, 300))
    gd, gc = gather_detections(det, cnt, total=total)   # ragged shards: sizes follow shard_bounds, nothing is exchanged
    # the handle form used by bench.py: start now, consume later; a second exchange may start before the first is read
    from unina_yolo_dla_b200.dp import DetectionGather
    G = DetectionGather()
    h1 = G.start(det, cnt, total)
    h2 = G.start(det.flip(0), cnt.flip(0), total)
    d1, c1 = h1.wait()
    d2, c2 = h2.wait()
    assert torch.equal(d1, gd) and torch.equal(c1, gc) and c2.shape == gc.shape and not torch.equal(d2, d1)
    if rank == 0:
        torch.save((gd, gc), out_path)
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for total in (0, 1, 5, 64, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_two_rank_gloo_gather_equals_single_process(tmp_path):
    from oracle import postproc as pp
    import uyd_testlib_cpu as T

    total, world = 5, 2          # uneven shards: 3 + 2 frames
    out = tmp_path / "gathered.pt"
    mp.spawn(_worker, args=(world, _free_port(), total, str(out)), nprocs=world, join=True)
    gd, gc = torch.load(out)
    y = T.synth_predictions(total, 4, 2100, seed=3, frac_conf=0.3)
    det, cnt = _pack(pp.non_max_suppression(y, 0.25, 0.7, 300))
    assert torch.equal(gc, cnt) and torch.equal(gd, det)
    assert int(cnt.sum()) > 0


def _eval_worker(rank, world, port, out_path):
    from unina_yolo_dla_b200.evaluate import DetectionEvaluator

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ev = DetectionEvaluator(device="cpu")
    ev.counters += torch.tensor([3 + rank, 1, 2 * rank])                      # what two shards would have counted
    ev._scores = [torch.tensor([[0.1 * (rank + 1), -1.0, 0.3], [-1.0, 0.2, -1.0]][: 2 - rank])]
    ev.reduce()
    if rank == 0:
        torch.save((ev.counters, ev.nonconformity_scores().sort().values), out_path)
    dist.destroy_process_group()


def test_two_rank_gloo_evaluator_reduce(tmp_path):
    """Per-rank evaluation + reduction: counters add up, the nonconformity scores of all ranks are pooled."""
    out = tmp_path / "eval.pt"
    mp.spawn(_eval_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    counters, scores = torch.load(out)
    assert counters.tolist() == [7, 2, 2]
    assert torch.allclose(scores, torch.tensor([0.1, 0.2, 0.2, 0.3, 0.3]))
