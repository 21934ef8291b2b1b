"""INT8 path: conv outputs bit-exact vs the integer fake-quant reference (oracle/quant.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [
    # cin, cout, k, stride, H, W, out dtype ("s8" | "f32" | "bf16"), impl
    (64, 64, 3, 1, 40, 40, "s8", "tc"),
    (64, 64, 3, 1, 32, 24, "f32", "tc"),
    (128, 64, 3, 1, 24, 24, "s8", "tc"),
    (32, 64, 3, 1, 32, 32, "s8", "tc"),
    (64, 64, 1, 1, 24, 24, "s8", "tc"),
    (256, 128, 1, 1, 16, 16, "s8", "tc"),
    (96, 32, 1, 1, 16, 16, "bf16", "tc"),
    (64, 64, 1, 1, 8, 8, "f32", "tc"),
    (32, 64, 3, 2, 32, 32, "s8", "tc"),
    (64, 128, 3, 2, 32, 32, "s8", "tc"),
    (16, 16, 3, 1, 20, 20, "s8", "direct"),
    (8, 8, 3, 1, 20, 20, "s8", "direct"),
    (16, 8, 1, 1, 16, 16, "f32", "direct"),
    (64, 64, 3, 1, 16, 16, "s8", "direct"),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-{c[1]}-k{c[2]}s{c[3]}-{c[6]}-{c[7]}")
def test_int8_conv_bit_exact(case):
    import unina_yolo_dla_b200 as uyd
    from unina_yolo_dla_b200._lib import IMPL_DIRECT, IMPL_TC, UYD_BF16, UYD_F32, UYD_S8
    from oracle import quant as oq

    cin, cout, k, stride, H, W, okind, impl = case
    rng = np.random.default_rng(11)
    B = 3
    x = rng.normal(0, 1.0, (B, cin, H, W)).astype(np.float32)
    w = (rng.normal(0, 1.0, (cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32)
    amax_x, amax_w = float(np.abs(x).max()), float(np.abs(w).max())
    qx, qw = oq.quantize(x, amax_x), oq.quantize(w, amax_w)
    gamma, beta = rng.uniform(0.8, 1.2, cout), rng.uniform(-0.1, 0.1, cout)
    mean, var = rng.normal(0, 0.2, cout), rng.uniform(0.8, 1.2, cout)
    mult, bias = oq.fold_multiplier(amax_x, amax_w, gamma, beta, mean, var, eps=1e-3)
    amax_next = 3.0
    out_scale = float(oq.scale_of(amax_next))
    acc, y, qy = oq.conv_int8(qx, qw, mult, bias, stride, relu=True, out_scale=out_scale)

    oh, ow = (H - 1) // stride + 1, (W - 1) // stride + 1
    p = uyd.Plan(0, B)
    src = p.buffer(H, W, cin, UYD_S8)
    dst = p.buffer(oh, ow, cout, {"s8": UYD_S8, "f32": UYD_F32, "bf16": UYD_BF16}[okind])
    p.conv_s8(src, dst, qw, mult, bias, k, stride, relu=True, out_scale=out_scale,
              impl=IMPL_TC if impl == "tc" else IMPL_DIRECT)
    p.finalize()
    p.write(src, torch.from_numpy(qx))
    p.run_no_input(B)
    torch.cuda.synchronize()
    got = p.read(dst, B).cpu()
    if okind == "s8":
        np.testing.assert_array_equal(got.numpy(), qy)
    elif okind == "f32":
        assert got.numpy().tobytes() == y.tobytes()          # fp32 requant epilogue: bit-exact
    else:
        want = torch.from_numpy(y).to(torch.bfloat16).float()
        assert torch.equal(got, want)
    assert np.abs(acc).max() > 1000                           # the accumulators are not trivially small


def test_int8_accumulator_beyond_2_pow_24_rounds_like_numpy():
    """|acc| > 2^24: the int32 -> fp32 conversion must round to nearest-even on both sides."""
    import unina_yolo_dla_b200 as uyd
    from unina_yolo_dla_b200._lib import IMPL_TC, UYD_F32, UYD_S8
    from oracle import quant as oq

    cin, cout, k = 128, 64, 3
    B, H, W = 1, 16, 16
    rng = np.random.default_rng(5)
    qx = np.full((B, cin, H, W), 127, np.int8)
    qx[:, ::7] = 126
    qw = rng.integers(120, 128, (cout, cin, k, k)).astype(np.int8)
    mult = np.full(cout, 1.0, np.float32) * np.float32(1.0000001)
    bias = rng.normal(0, 1, cout).astype(np.float32)
    acc, y, _ = oq.conv_int8(qx, qw, mult, bias, 1, relu=False)
    assert np.abs(acc).max() > 2 ** 24
    p = uyd.Plan(0, B)
    src = p.buffer(H, W, cin, UYD_S8)
    dst = p.buffer(H, W, cout, UYD_F32)
    p.conv_s8(src, dst, qw, mult, bias, k, 1, relu=False, impl=IMPL_TC)
    p.finalize()
    p.write(src, torch.from_numpy(qx))
    p.run_no_input(B)
    torch.cuda.synchronize()
    assert p.read(dst, B).cpu().numpy().tobytes() == y.tobytes()


def test_quantize_op_and_residual_and_depthwise_are_bit_exact():
    """The pieces the INT8 network adds to the single conv: the input quantiser (bf16 -> int8), the
    residual add after the activation (tensor-core and dp4a kernels) and the depth-wise int8 conv."""
    import unina_yolo_dla_b200 as uyd
    from unina_yolo_dla_b200._lib import IMPL_DIRECT, IMPL_TC, UYD_BF16, UYD_S8
    from oracle import quant as oq
    from oracle.quant_graph import bf16

    rng = np.random.default_rng(3)
    B, H, W = 2, 24, 40
    for cin, cout, k, impl, dw in ((64, 64, 3, IMPL_TC, False), (32, 32, 1, IMPL_TC, False), (8, 8, 3, IMPL_DIRECT, False),
                                   (4, 4, 3, IMPL_DIRECT, False), (32, 32, 3, IMPL_DIRECT, True), (128, 128, 3, IMPL_DIRECT, True)):
        x = bf16(torch.from_numpy(rng.normal(0, 1, (B, cin, H, W)).astype(np.float32)))
        res = bf16(torch.from_numpy(rng.normal(0, 1, (B, cout, H, W)).astype(np.float32)))
        w = (rng.normal(0, 1, (cout, 1 if dw else cin, k, k)) / np.sqrt((1 if dw else cin) * k * k)).astype(np.float32)
        ax, aw = float(x.abs().max()), float(np.abs(w).max())
        qx, qw = oq.quantize(x.numpy(), ax), oq.quantize(w, aw)
        mult, bias = oq.fold_multiplier(ax, aw, rng.uniform(0.8, 1.2, cout), rng.uniform(-0.1, 0.1, cout),
                                        rng.normal(0, 0.2, cout), rng.uniform(0.8, 1.2, cout), eps=1e-3)
        _, y, _ = oq.conv_int8(qx, qw, mult, bias, 1, relu=True, groups=cin if dw else 1)
        use_res = not dw
        want = bf16(torch.from_numpy(y) + res) if use_res else bf16(torch.from_numpy(y))
        p = uyd.Plan(0, B)
        src = p.buffer(H, W, cin + 8).sub(8, cin)            # bf16 activation in a channel slice
        q = p.buffer(H, W, cin, UYD_S8)
        dst = p.buffer(H, W, cout)
        rs = p.buffer(H, W, cout + 16).sub(16, cout)
        p.quantize(src, q, float(oq.scale_of(ax)))
        p.conv_s8(q, dst, qw, mult, bias, k, 1, relu=True, impl=impl, depthwise=dw, res=rs if use_res else None)
        p.finalize()
        p.write(src, x)
        p.write(rs, res)
        p.run_no_input(B)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(p.read(q, B).cpu().numpy(), qx)
        assert torch.equal(p.read(dst, B).cpu(), want), (cin, cout, k, dw)


def test_int8_network_is_bit_exact_from_the_last_float_layer():
    """BASELINE config 3: the whole INT8 graph (static max-calibrated scales, model.0-2 float) reproduces the
    integer reference byte for byte on the raw head outputs, given the tensor the float layers produce; the
    decoded prediction (DFL projection quantised too) agrees within the exp() tolerance and NMS on it is exact."""
    import unina_yolo_dla_b200 as uyd
    from oracle import init as oi
    from oracle import postproc as pp
    from oracle import yolo_graph as yg
    from oracle.quant_graph import Int8Graph

    m = uyd.UninaYoloB200.from_yaml().init_synthetic(seed=0).cuda()
    x = oi.seeded_frames(2, 320, seed=11).cuda()
    m.calibrate_cls_bias(x, 500, 0.25)
    amax = m.calibrate_int8(x)
    # 157 convs (the stem reads the frame and stays float) + the DFL projection
    assert len(amax) == 158 and "model.0.conv" not in amax and "model.20.dfl.conv" in amax and m.quant is not None
    y, raws = m(x)
    torch.cuda.synchronize()
    p = m.plan_for(x)
    texts = [p.op_info(i)[0] for i in range(p.launches)]
    n_fold = sum(1 for t in texts if t.startswith("conv_s8") and "+up(partial)" in t)   # a folded Upsample + Concat + cv1 = two launches
    n_s8 = sum(1 for t in texts if t.startswith("conv_s8")) - n_fold + 7 * sum(1 for t in texts if t.startswith("c3k_fused_s8"))
    assert n_s8 == 158 - 2 - 16                               # every conv outside model.0, model.1 and model.2 (16 convs)
    assert n_fold == 2 and not any(t.startswith("upsample2x") for t in texts)   # both Upsample layers are folded into their consumer
    assert sum(1 for t in texts if t.startswith("c3k_fused_s8")) == 9   # the 80 x 80 and 40 x 40 C3k blocks run as one launch each
    l2 = p.read(p.layer_outputs[2], 2).cpu()
    ref = yg.DetectionModel(yg.default_yaml_path())
    ref.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()}, strict=True)
    ref.eval()
    g = Int8Graph(ref, amax)
    want = g.forward_from({2: l2})
    for a, b in zip(raws, want):
        assert a.shape == b.shape and a.cpu().numpy().tobytes() == b.numpy().tobytes()
    y_ref = g.decode(ref.model[-1], want, "model.20.dfl.conv")
    assert float((y.cpu()[:, :4] - y_ref[:, :4]).abs().max()) < 0.6     # one quantisation step of a probability ~ 0.5 px at stride 16
    assert float((y.cpu()[:, 4:] - y_ref[:, 4:]).abs().max()) < 2e-6
    det, cnt, idx = m.nms(y, 0.25, 0.7, 300, return_index=True)
    wdet, widx = pp.non_max_suppression(y.cpu().numpy(), 0.25, 0.7, 300, return_index=True)
    for b in range(2):
        n = int(cnt[b])
        assert n == len(wdet[b]) > 0 and np.array_equal(idx[b, :n].cpu().numpy(), widx[b])
    # the INT8 prediction stays close to the bf16 one (sanity of the scales, not a parity claim)
    m.set_quantization(None)
    y_f = m(x, raw_heads=False)
    assert float((y.cpu()[:, 4:] - y_f.cpu()[:, 4:]).abs().max()) < 0.25
    # a QAT checkpoint (reference schema + _amax buffers) switches the INT8 path on again
    sd = dict(m.state_dict())
    sd.update({k: v for k, v in uyd.UninaYoloB200.from_yaml().set_quantization(amax).quant_state_dict().items()})
    m2 = uyd.UninaYoloB200.from_yaml()
    m2.load_state_dict(sd)
    assert m2.quant is not None and m2.quant.amax.keys() == amax.keys()


@pytest.mark.parametrize("c,H,W,B", [(8, 40, 80, 3), (16, 40, 40, 2), (32, 20, 40, 2), (8, 160, 160, 2), (16, 80, 80, 5), (32, 40, 40, 9)])
def test_fused_int8_c3k_block_equals_the_unfused_ops(c, H, W, B, monkeypatch):
    """uyd_plan_add_c3k_s8 (one launch: int8 codes as exact bf16 values on the flat-frame mma.sync kernel, requantisation
    in the stage epilogues) against the seven conv_s8 + quantize ops the same C3k emits when the fusion is switched off:
    the bf16 output bytes must be identical (the unfused ops are the ones pinned to the integer oracle above).  Channel
    slices of wider buffers, tiles at every image border, several tiles per image."""
    import torch.nn as nn
    import unina_yolo_dla_b200 as uyd
    from unina_yolo_dla_b200 import quant as Q
    from unina_yolo_dla_b200 import yolo as Y

    g = torch.Generator().manual_seed(100 * c + H)
    blk = Y.C3k(c, c, 2)
    amax = {}
    for name, mod in blk.named_modules():
        if isinstance(mod, nn.Conv2d):
            mod._uyd_name = "model.9." + name
            mod.weight.data = torch.randn(mod.weight.shape, generator=g) / (mod.weight[0].numel() ** 0.5)
            amax[mod._uyd_name] = (float(torch.empty(1).uniform_(2.0, 5.0, generator=g)), float(mod.weight.detach().abs().max()))
        if isinstance(mod, nn.BatchNorm2d):
            mod.weight.data = torch.empty(mod.num_features).uniform_(0.6, 1.4, generator=g)
            mod.bias.data = torch.randn(mod.num_features, generator=g) * 0.2
            mod.running_mean = torch.randn(mod.num_features, generator=g) * 0.2
            mod.running_var = torch.empty(mod.num_features).uniform_(0.5, 1.5, generator=g)
    blk.eval()
    x = (torch.randn(B, c, H, W, generator=g) * 1.5).relu()

    def run(fused: bool):
        monkeypatch.setenv("UYD_INT8_NO_C3K_FUSION", "0" if fused else "1")
        p = uyd.Plan(0, B)
        p.quant = Q.QuantSpec(dict(amax), ())
        p.fusion = True
        src = p.buffer(H, W, 3 * c).sub(c, c)
        dst = p.buffer(H, W, 2 * c).sub(c, c)
        blk.emit(p, src, dst)
        p.finalize()
        p.write(src, x)
        p.run_no_input(B)
        torch.cuda.synchronize()
        texts = [p.op_info(i)[0] for i in range(p.launches)]
        return p.read(dst, B).cpu(), texts, float(p.read(p.buffer_slice(dst.buf, 0, c), B).abs().max())

    got, t_fused, outside = run(True)
    want, t_unfused, _ = run(False)
    assert len(t_fused) == 1 and t_fused[0].startswith("c3k_fused_s8") and len(t_unfused) >= 11
    assert outside == 0.0
    assert float(want.abs().max()) > 0.1
    assert got.numpy().tobytes() == want.numpy().tobytes()


def test_int8_predict_is_batch_independent_and_survives_graph_replay():
    """The INT8 plan (fused C3k blocks, folded Upsample + Concat) through predict_batched: a batch of 2 takes the
    CUDA-graph replay path, a batch of 12 the stream path with different tile heights and grid sizes -- the detections
    of the same frames must be identical bytes (no batch coupling, no dependence on the tiling)."""
    import unina_yolo_dla_b200 as uyd

    m = uyd.UninaYoloB200.from_yaml().init_synthetic(0).cuda()
    x = torch.rand(12, 3, 640, 640, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    m.calibrate_cls_bias(x[:8], 300, 0.25)
    m.calibrate_int8(x[:8])
    d12, c12 = (t.clone() for t in m.predict_batched(x, 0.25, 0.7, 300))
    for _ in range(3):
        d2, c2 = m.predict_batched(x[:2].contiguous(), 0.25, 0.7, 300)
    torch.cuda.synchronize()
    assert int(c2.min()) > 0 and torch.equal(c2, c12[:2]) and torch.equal(d2, d12[:2])
    p = m.plan_for(x)
    assert sum(1 for i in range(p.launches) if p.op_info(i)[0].startswith("c3k_fused_s8")) == 14
