"""CPU tests of the host side: the plan the product emits computes the oracle's network
(checked by executing the emitted op list with torch), state_dict schema parity, loud failure
without a GPU, and the C-ABI library exporting every symbol of include/uyd.h."""
import re
from pathlib import Path

import pytest
import torch

import unina_yolo_dla_b200 as uyd
from fake_plan import FakePlan
from oracle import init as oi
from oracle import yolo_graph as yg

ROOT = Path(__file__).resolve().parents[1]


def _emit_fake(model, B, H, W, monkeypatch, fused=False):
    import unina_yolo_dla_b200.yolo as ymod

    monkeypatch.setattr(ymod, "Plan", FakePlan)
    return model._build_plan(0, B, H, W, fused)


def test_emitted_plan_equals_oracle_network(monkeypatch):
    m = uyd.UninaYoloB200.from_yaml().init_synthetic(seed=3)
    ref = yg.DetectionModel(yg.default_yaml_path())
    ref.load_state_dict(m.state_dict(), strict=True)
    ref.eval()
    x = oi.seeded_frames(2, 128, seed=5)
    with torch.no_grad():
        (y_ref, raw_ref), feats = ref.forward_features(x, want=set(range(20)))
    p = _emit_fake(m, 2, 128, 128, monkeypatch)
    bufs = p.execute(x)
    for i, s in enumerate(p.layer_outputs):
        if s is None:
            continue
        got = bufs[s.buf][:, s.coff:s.coff + s.c]
        assert torch.allclose(got, feats[i], rtol=1e-4, atol=1e-4), f"layer {i} differs"
    for h, r in zip(p.heads, raw_ref):
        assert torch.allclose(bufs[h.buf], r, rtol=1e-4, atol=1e-4)
    # census: 158 convs (3 inside the fused stem: layers 0, 1 and model.2.cv1; 19 inside the 8 chained head launches)
    # + 1 pool cascade + 2 upsamples, no concat / chunk op at all
    kinds = [o[0] for o in p.ops]
    assert kinds.count("stem2") == 1 and kinds.count("chain") == 8 and kinds.count("conv") == 158 - 3 - 19 + 2
    # + 2 half-resolution partial-sum convs: both Upsample + Concat pairs are folded into the next C3k2's first conv
    assert kinds.count("sppf") == 1 and kinds.count("up") == 0


def test_unchained_plan_is_the_plain_conv_list(monkeypatch):
    monkeypatch.setenv("UYD_NO_CHAIN", "1")
    monkeypatch.setenv("UYD_NO_STEM_FUSION", "1")
    monkeypatch.setenv("UYD_NO_UPSAMPLE_FOLD", "1")
    m = uyd.UninaYoloB200.from_yaml().init_synthetic(seed=3)
    p = _emit_fake(m, 1, 128, 128, monkeypatch)
    kinds = [o[0] for o in p.ops]
    assert kinds.count("conv") == 158 and kinds.count("chain") == 0 and kinds.count("up") == 2


def test_fused_decode_plan_writes_the_oracle_prediction(monkeypatch):
    """The plan variant whose head kernels decode in their epilogue has no head buffers and its
    emitted ops compute the oracle's y[B, 4+nc, A] (DFL decode + sigmoid)."""
    m = uyd.UninaYoloB200.from_yaml().init_synthetic(seed=5)
    ref = yg.DetectionModel(yg.default_yaml_path())
    ref.load_state_dict(m.state_dict(), strict=True)
    ref.eval()
    x = oi.seeded_frames(2, 128, seed=7)
    with torch.no_grad():
        y_ref, _ = ref(x)
    p = _emit_fake(m, 2, 128, 128, monkeypatch, fused=True)
    assert p.fused and all(h.buf < 0 for h in p.heads)
    p.execute(x)
    assert p.y.shape == y_ref.shape
    assert torch.allclose(p.y[:, :4], y_ref[:, :4], rtol=1e-4, atol=1e-3)
    assert torch.allclose(p.y[:, 4:], y_ref[:, 4:], rtol=1e-4, atol=1e-5)


def test_state_dict_schema_matches_oracle_and_roundtrips():
    m = uyd.UninaYoloB200.from_yaml()
    o = yg.DetectionModel(yg.default_yaml_path())
    a, b = m.state_dict(), o.state_dict()
    assert list(a.keys()) == list(b.keys()) and len(a) == 925
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    m.load_state_dict(b, strict=True)
    o.load_state_dict(m.state_dict(), strict=True)
    det = m.model[-1]
    assert (det.nl, det.nc, det.reg_max, det.no) == (3, 4, 16, 68)
    assert det.stride.tolist() == [4.0, 8.0, 16.0] and len(det.cv2) == 3 and len(det.cv3) == 3
    assert m.names == {0: "0", 1: "1", 2: "2", 3: "3"} and m.nc == 4


def test_product_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = uyd.UninaYoloB200.from_yaml()
    with pytest.raises(uyd.UydError):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(uyd.UydError):
        uyd.Plan(0, 1)
    m.train()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64))


def test_library_exports_every_declared_symbol():
    from unina_yolo_dla_b200 import _lib

    header = (ROOT / "include" / "uyd.h").read_text()
    declared = set(re.findall(r"\b(uyd_[a-z0-9_]+)\s*\(", header))
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"libuyd.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert L.uyd_version() == 100


def test_product_does_not_import_the_oracle():
    for f in (ROOT / "unina-yolo-dla_b200").glob("*.py"):
        assert "oracle" not in f.read_text(), f"{f.name} mentions the oracle"


def test_emitted_plan_uses_fused_c3k_at_full_resolution(monkeypatch):
    """At 640x640 every C3k interior (16 blocks x 7 convs) is emitted as one fused op and the plan
    still computes the oracle's network."""
    m = uyd.UninaYoloB200.from_yaml().init_synthetic(seed=4)
    ref = yg.DetectionModel(yg.default_yaml_path())
    ref.load_state_dict(m.state_dict(), strict=True)
    ref.eval()
    x = oi.seeded_frames(1, 640, seed=6)
    with torch.no_grad():
        y_ref, raw_ref = ref(x)
    p = _emit_fake(m, 1, 640, 640, monkeypatch)
    kinds = [o[0] for o in p.ops]
    # the fused stem (3 convs), 16 fused C3k blocks (7 convs each) and 8 chained head launches (19 convs)
    assert kinds.count("stem2") == 1 and kinds.count("c3k") == 16 and kinds.count("chain") == 8
    assert kinds.count("conv") == 158 - 3 - 16 * 7 - 19 + 2 and kinds.count("up") == 0
    bufs = p.execute(x)
    for h, r in zip(p.heads, raw_ref):
        assert torch.allclose(bufs[h.buf], r, rtol=1e-4, atol=1e-4)
