"""Host-to-device bandwidth of all ranks at once (torchrun): pinned vs write-combined staging, one or two copy streams,
whole batch vs 4 chunks.  Answers which host-side setting bounds e2e at N = 8 (bench.py e2e.bound_probe shows THAT it does).
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/h2d_probe.py"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from unina_yolo_dla_b200.hostmem import pinned_frames  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, S, STEPS = 64, 640, 20
shape = (B, 3, S, S)
dev_buf = [torch.empty(shape, dtype=torch.uint8, device="cuda") for _ in range(2)]


def measure(host, streams, chunks):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    def go():
        for i in range(STEPS):
            s = ss[i % streams]
            with torch.cuda.stream(s):
                for c in range(chunks):
                    lo, hi = c * B // chunks, (c + 1) * B // chunks
                    dev_buf[i % 2][lo:hi].copy_(host[lo:hi], non_blocking=True)
    go()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in ss:
        s.wait_stream(torch.cuda.current_stream())
    go()
    for s in ss:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return host.numel() * STEPS / (float(t.item()) * 1e-3) / 1e9


for name, host in (("pinned", torch.empty(shape, dtype=torch.uint8).pin_memory()), ("write-combined", pinned_frames(shape, torch.uint8, True))):
    host.fill_(7) if name == "pinned" else host.view(-1)[:: 4096].fill_(7)
    for streams, chunks in ((1, 1), (2, 1), (1, 4)):
        g = measure(host, streams, chunks)
        if rank == 0:
            print(f"N={world} {name:15s} streams {streams} chunks {chunks}: {g:6.1f} GB/s per GPU (slowest rank), {g * world:7.1f} GB/s total", flush=True)
if world > 1:
    dist.destroy_process_group()
