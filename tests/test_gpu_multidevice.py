"""Two GPUs driven by ONE process (include/uyd.h advertises a per-device handle): plans on cuda:0 and cuda:1 give the
same bytes, no entry point changes the caller's current device, and the > 48 KB shared-memory opt-in reaches every
device.  Needs two visible GPUs (gpurun --gpus 2); skipped otherwise."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_devices_in_one_process_agree():
    import unina_yolo_dla_b200 as uyd
    from oracle import init as oi

    x = oi.seeded_frames(3, 320, seed=31)
    torch.cuda.set_device(0)
    outs = []
    for dev in (1, 0, 1):   # the first launches of every kernel family happen on device 1, not on the current device
        m = uyd.UninaYoloB200.from_yaml().init_synthetic(seed=0).to(f"cuda:{dev}")
        xd = x.to(f"cuda:{dev}")
        y = m(xd, raw_heads=False)
        det, cnt = m.nms(y, 0.05, 0.7, 300)
        c = uyd.UninaCustomB200(4, 8).init_synthetic(seed=1).to(f"cuda:{dev}")
        heads = c(xd)
        torch.cuda.synchronize(dev)
        assert torch.cuda.current_device() == 0, "an entry point changed the caller's current device"
        assert y.device.index == dev and det.device.index == dev
        outs.append((y.cpu(), det.cpu(), cnt.cpu(), [t.cpu() for pair in heads for t in pair]))
        del m, c   # plan destruction (uyd_plan_destroy) must not move the current device either
        assert torch.cuda.current_device() == 0
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) and torch.equal(o[2], outs[0][2])
        for a, b in zip(o[3], outs[0][3]):
            assert torch.equal(a, b)
