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
