"""GPU parity of the custom variant (model.py network, TLBR decode, postprocess.hpp NMS)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _decode_gpu(cls, reg, stride, thr, q, strict=1, base=0):
    from unina_yolo_dla_b200 import _lib

    nc, h, w = cls.shape
    cap = h * w
    dets = torch.zeros(cap, 8, device="cuda")
    cell = torch.zeros(cap, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    c, r = torch.from_numpy(cls).cuda(), torch.from_numpy(reg).cuda()
    _lib.check(_lib.lib().uyd_decode_tlbr(_lib.context(0), C.c_void_p(c.data_ptr()), C.c_void_p(r.data_ptr()),
                                          C.c_void_p(dets.data_ptr()), C.c_void_p(cell.data_ptr()), C.c_void_p(cnt.data_ptr()),
                                          cap, w, h, stride, nc, thr, q, strict, base, None))
    torch.cuda.synchronize()
    n = int(cnt.item())
    return dets[:n], cell[:n], cnt, dets, cell


def _as_records(t):
    from oracle import postproc as pp

    a = t.cpu().numpy()
    out = np.zeros(len(a), dtype=pp.DET_DTYPE)
    for i, f in enumerate(("x1", "y1", "x2", "y2", "conf")):
        out[f] = a[:, i]
    out["cls"] = a[:, 5].view(np.int32)
    return out


@pytest.mark.parametrize("q", [0.0, 0.1])
def test_tlbr_decode_and_hpp_nms_match_oracle(q):
    from unina_yolo_dla_b200 import _lib
    from oracle import postproc as pp

    rng = np.random.default_rng(7)
    nc, h, w, stride, thr = 4, 80, 80, 8, 0.5
    cls = rng.normal(-0.3, 1.5, (nc, h, w)).astype(np.float32)
    reg = rng.uniform(0.5, 4.0, (4, h, w)).astype(np.float32)
    want = pp.decode_tlbr(cls, reg, stride, thr, q)
    got, cell, cnt, dets_full, cell_full = _decode_gpu(cls, reg, stride, thr, q)
    order = torch.argsort(cell)
    rec = _as_records(got[order])
    cells = cell[order].cpu().numpy()
    conf_all = (1.0 / (1.0 + np.exp(-cls.astype(np.float64)))).max(0).reshape(-1)
    want2, want_cells = pp.decode_tlbr(cls, reg, stride, thr, q, return_cells=True)
    assert want2.tobytes() == want.tobytes()
    assert len(want_cells) == len(want)
    # compare the INTERSECTION by cell index; a cell may be on one side only iff its score is within 2 ulp of thr
    common, gi, wi = np.intersect1d(cells, want_cells, return_indices=True)
    only = np.setxor1d(cells, want_cells)
    assert len(only) <= 2 and np.all(np.abs(conf_all[only] - thr) <= 3e-7), (len(only), conf_all[only])
    assert len(common) >= len(want) - 2
    for f in ("x1", "y1", "x2", "y2"):
        np.testing.assert_array_equal(rec[f][gi], want[f][wi])      # boxes: bit-exact fp32
    np.testing.assert_array_equal(rec["cls"][gi], want["cls"][wi])
    np.testing.assert_allclose(rec["conf"][gi], want["conf"][wi], rtol=0, atol=3e-7)
    # NMS on the GPU-decoded detections vs the oracle's postprocess.hpp statement: byte-equal
    L = _lib.lib()
    cap = h * w
    ws_bytes = int(L.uyd_nms_detections_workspace_bytes(cap))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    kept = torch.zeros(1024, 8, device="cuda")
    kcnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(L.uyd_nms_detections(_lib.context(0), C.c_void_p(dets_full.data_ptr()), C.c_void_p(cell_full.data_ptr()),
                                    C.c_void_p(cnt.data_ptr()), cap, 0.45, C.c_void_p(ws.data_ptr()), ws_bytes,
                                    C.c_void_p(kept.data_ptr()), C.c_void_p(kcnt.data_ptr()), None))
    torch.cuda.synchronize()
    k = _as_records(kept[: int(kcnt.item())])
    # oracle order: confidence descending, ties by cell index (records sorted by cell first)
    want_k = pp.greedy_nms_hpp(rec, 0.45)[:1024]
    assert len(k) == len(want_k)
    assert k.tobytes() == want_k.tobytes()
    if pp.ref_lib() is not None and len(np.unique(rec["conf"])) == len(rec):
        assert pp.greedy_nms_hpp(rec, 0.45, use_ref=True)[:1024].tobytes() == k.tobytes()


@pytest.mark.parametrize("bc", [8, 32])
def test_custom_forward_matches_oracle(bc):
    import unina_yolo_dla_b200 as uyd
    from oracle import custom_graph as cg
    from oracle import init as oi

    m = uyd.UninaCustomB200(4, bc).init_synthetic(seed=1)
    ref = cg.CustomNet(4, bc)
    ref.load_state_dict(m.state_dict(), strict=True)
    ref.eval()
    m = m.cuda()
    x = oi.seeded_frames(2, 320, seed=9)
    with torch.no_grad():
        want = ref(x)
    got = m(x.cuda())
    torch.cuda.synchronize()
    for (gc, gr), (wc, wr) in zip(got, want):
        assert gc.shape == wc.shape and gr.shape == wr.shape
        assert float((gc.cpu() - wc).abs().max() / wc.abs().max()) <= 1e-2
        assert float((gr.cpu() - wr).abs().max() / wr.abs().max()) <= 1e-2


def test_custom_forward_640_matches_pinned_oracle():
    """The benched configuration of the model.py variant (base_channels 32, 640 x 640; custom_variant in bench.py) against
    the oracle that is pinned bit-exactly to the real model.py.  Measured on B200 (tools/custom_parity_diag.py, four
    weight seeds x two extents): regression maps 1.1e-3 .. 1.9e-3 of their range, class maps 5e-3 .. 1.33e-2 (rms
    1.3e-3 .. 3.1e-3) -- bf16 activations through 41 sequential convs onto a class map whose own range is small; the
    same figures with the fp32-input CUDA-core stem, so the split-bf16 stem adds nothing.  Stated tolerance for this
    variant: regression <= 5e-3, class <= 2e-2 of range and <= 5e-3 rms (the 1e-2 of the north star holds for the YAML
    network, tests/test_gpu_northstar.py, and for this variant's seeds 1 / 320 x 320 case above)."""
    import unina_yolo_dla_b200 as uyd
    from oracle import custom_graph as cg
    from oracle import init as oi

    m = uyd.UninaCustomB200(4, 32).init_synthetic(seed=3)
    ref = cg.CustomNet(4, 32)
    ref.load_state_dict(m.state_dict(), strict=True)
    ref.eval()
    m = m.cuda()
    x = oi.seeded_frames(4, 640, seed=11)
    with torch.no_grad():
        want = ref(x)
    got = m(x.cuda())
    torch.cuda.synchronize()
    worst = {"cls": 0.0, "reg": 0.0, "cls_rms": 0.0}
    for pair_g, pair_w in zip(got, want):
        for name, g, w in zip(("cls", "reg"), pair_g, pair_w):
            assert g.shape == w.shape
            d = (g.cpu() - w).abs()
            rng = float(w.abs().max())
            worst[name] = max(worst[name], float(d.max()) / rng)
            if name == "cls":
                worst["cls_rms"] = max(worst["cls_rms"], float(d.pow(2).mean().sqrt()) / rng)
    print(f"model.py variant bc32 640x640 batch 4 (range-normalised): class {worst['cls']:.3e} (rms {worst['cls_rms']:.3e}), regression {worst['reg']:.3e}")
    assert worst["reg"] <= 5e-3 and worst["cls"] <= 2e-2 and worst["cls_rms"] <= 5e-3


def test_custom_predict_rows():
    import unina_yolo_dla_b200 as uyd
    from oracle import init as oi

    m = uyd.UninaCustomB200(4, 8).init_synthetic(seed=1).cuda()
    x = oi.seeded_frames(2, 256, seed=9).cuda()
    res = m.predict(x, conf=0.3, iou=0.45)
    assert len(res) == 2
    for r in res:
        assert r.ndim == 2 and r.shape[1] == 6 and r.shape[0] <= 1024
        if len(r):
            assert bool((r[:-1, 4] >= r[1:, 4]).all())  # kept order = confidence order
            assert set(r[:, 5].tolist()) <= {0.0, 1.0, 2.0, 3.0}


@pytest.mark.parametrize("n", [1, 37, 700, 1024])
def test_inplace_nms_leaves_the_buffer_like_run_gpu_nms(n):
    """uyd_nms_detections_inplace (what libuyd_compat's run_gpu_nms calls): the first n records sorted by
    confidence, valid = survivors of the exact greedy NMS of postprocess.hpp; uyd_compact_valid = the ordered
    compaction of copy_valid_detections_to_host."""
    from unina_yolo_dla_b200 import _lib
    from oracle import postproc as pp

    rng = np.random.default_rng(n)
    rec = np.zeros(n, dtype=pp.DET_DTYPE)
    cx, cy = rng.uniform(20, 300, n), rng.uniform(20, 300, n)
    w, h = rng.uniform(10, 80, n), rng.uniform(10, 80, n)
    rec["x1"], rec["y1"], rec["x2"], rec["y2"] = cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2
    rec["conf"] = rng.permutation(n).astype(np.float32) / n * 0.5 + 0.5   # distinct
    rec["cls"] = rng.integers(0, 4, n)
    raw = np.zeros((1024, 8), np.float32)
    for i, f in enumerate(("x1", "y1", "x2", "y2", "conf")):
        raw[:n, i] = rec[f]
    raw[:n, 5] = rec["cls"].astype(np.int32).view(np.float32)
    raw[:n, 6] = np.int32(1).view(np.float32)
    raw[n:, 4] = 2.0  # records past n must not be touched
    dets = torch.from_numpy(raw).cuda()
    kcnt = torch.zeros(2, dtype=torch.int32, device="cuda")
    L = _lib.lib()
    _lib.check(L.uyd_nms_detections_inplace(_lib.context(0), C.c_void_p(dets.data_ptr()), None, n, 0.45, C.c_void_p(kcnt.data_ptr()), None))
    out = torch.zeros(1024, 8, device="cuda")
    _lib.check(L.uyd_compact_valid(_lib.context(0), C.c_void_p(dets.data_ptr()), n, C.c_void_p(out.data_ptr()),
                                   C.c_void_p(kcnt[1:].data_ptr()), None))
    torch.cuda.synchronize()
    got = dets.cpu().numpy()
    assert np.array_equal(got[n:], raw[n:])
    order = np.argsort(-rec["conf"], kind="stable")
    np.testing.assert_array_equal(got[:n, :5], raw[:n][order][:, :5])          # sorted in place, bit-exact
    want_k = pp.greedy_nms_hpp(rec, 0.45)
    valid = got[:n, 6].view(np.int32)
    assert int(kcnt[0]) == len(want_k) == int(valid.sum()) == int(kcnt[1])
    k = _as_records(out[: len(want_k)])
    assert k.tobytes() == want_k.tobytes()
    if pp.ref_lib() is not None:
        assert pp.greedy_nms_hpp(rec, 0.45, use_ref=True).tobytes() == k.tobytes()
