"""f-4 histogram / entropy calibrator and the SmallObjectMetric consumer: CPU pins + GPU parity."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden" / "small_object_metric.npz"
REF = Path("/root/reference/unina_yolo_dla")


# ---------------------------------------------------------------- SmallObjectMetric (data_loader.py:249-414)
def _running(update, thr):
    d = np.load(GOLD)
    tot, run = np.zeros(3, np.int64), []
    for i in range(len(d["pred"])):
        tot += np.asarray(update([d["pred"][i, : d["npred"][i]]], [d["gt"][i, : d["ngt"][i]]], thr))
        run.append(tot.copy())
    return np.asarray(run), d


@pytest.mark.parametrize("thr", [0.5, 0.3])
def test_small_object_oracle_matches_reference_golden(thr):
    from oracle import evalref as er

    run, d = _running(lambda p, g, t: er.small_object_metric_update(p, g, 15, t, 640), thr)
    np.testing.assert_array_equal(run, d[f"running_thr{thr}"])
    assert run[-1].sum() > 40


@pytest.mark.skipif(not REF.exists(), reason="reference mount absent")
def test_small_object_oracle_matches_live_reference():
    sys.path.insert(0, str(REF))
    from data_loader import SmallObjectMetric as RefMetric
    from oracle import evalref as er

    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    import make_small_object_golden as mk

    pred, npred, gt, ngt = mk.cases(seed=5, n_img=16)
    m = RefMetric(15, 0.5, 640)
    P = [torch.from_numpy(pred[i, : npred[i]]) for i in range(len(pred))]
    G = [torch.from_numpy(gt[i, : ngt[i]]) for i in range(len(gt))]
    m.update(P, G)
    assert (m.true_positives, m.false_positives, m.false_negatives) == er.small_object_metric_update(
        [p.numpy() for p in P], [g.numpy() for g in G], 15, 0.5, 640)


@pytest.mark.gpu
@pytest.mark.parametrize("thr", [0.5, 0.3])
def test_small_object_metric_gpu_matches_reference_golden(thr):
    from unina_yolo_dla_b200.evaluate import SmallObjectMetric

    d = np.load(GOLD)
    m = SmallObjectMetric(15, thr, 640)
    P = [torch.from_numpy(d["pred"][i, : d["npred"][i]]) for i in range(len(d["pred"]))]
    G = [torch.from_numpy(d["gt"][i, : d["ngt"][i]]) for i in range(len(d["gt"]))]
    m.update(P[:10], G[:10])           # two batched updates accumulate like 24 single-image updates
    assert [m.true_positives, m.false_positives, m.false_negatives] == d[f"running_thr{thr}"][9].tolist()
    m.update(P[10:], G[10:])
    assert [m.true_positives, m.false_positives, m.false_negatives] == d[f"running_thr{thr}"][-1].tolist()
    c = m.compute()
    np.testing.assert_allclose([c["small_object_precision"], c["small_object_recall"], c["small_object_f1"]], d[f"prf_thr{thr}"], rtol=1e-12)
    # predict-format rows convert to the metric's format
    det = torch.tensor([[[100.0, 200.0, 110.0, 212.0, 0.9, 2.0]]])
    r = SmallObjectMetric.from_xyxy_pixels(det, 640.0)[0, 0]
    np.testing.assert_allclose(r.numpy(), [105 / 640, 206 / 640, 10 / 640, 12 / 640, 0.9, 2.0], rtol=1e-6)
    m.reset()
    assert m.compute()["small_object_tp"] == 0


# ---------------------------------------------------------------- histogram / entropy calibration (qat.py:91-126, 676-697)
def test_entropy_amax_equals_the_loop_form_restatement():
    import unina_yolo_dla_b200.quant as Q
    from oracle import quant as oq

    rng = np.random.default_rng(0)
    for k, (scale, outliers) in enumerate(((1.0, 0), (0.3, 4), (2.0, 30))):
        x1 = np.abs(rng.normal(0, scale, 60000)).astype(np.float32)
        x2 = np.abs(rng.normal(0, scale * 1.2, 30000)).astype(np.float32)
        x2[:outliers] *= 5                                   # a later batch exceeds the first range: the histogram grows
        c = Q.HistogramCalibrator(num_bins=512)
        c.collect_host(x1)
        c.collect_host(x2)
        h, e = oq.histogram_collect([x1, x2], num_bins=512)
        assert np.array_equal(c.hist, h) and np.allclose(c.edges, e) and h.sum() == 90000
        a = c.compute_amax("entropy")
        assert a == oq.amax_entropy_loops(h, e)
        assert 2.0 * scale < a <= e[-1]                      # clips the tail, keeps the bulk
        assert c.compute_amax("percentile", 99.9) < c.compute_amax("percentile", 99.999) <= e[-1]
        assert 0 < c.compute_amax("mse") <= e[-1]
        assert abs(c.compute_amax("max") - max(x1.max(), x2.max())) <= c.width * 1.001


@pytest.mark.gpu
def test_gpu_histogram_and_entropy_calibration():
    """uyd_plan_slice_histogram == the host binning rule on the same bf16 tensor; calibrate_int8(method=...) yields
    scales that clip (entropy <= max) and the INT8 graph stays bit-exact w.r.t. the integer oracle with them."""
    import unina_yolo_dla_b200 as uyd
    import unina_yolo_dla_b200.quant as Q
    from oracle import init as oi
    from oracle import yolo_graph as yg
    from oracle.quant_graph import Int8Graph

    g = torch.Generator().manual_seed(3)
    x = (torch.randn(3, 24, 20, 32, generator=g) * 1.7).to(torch.bfloat16).float()
    p = uyd.Plan(0, 3)
    buf = p.buffer(20, 32, 40)
    sl = buf.sub(8, 24)
    dummy = p.buffer(20, 32, 8)
    p.conv(buf.sub(0, 8), dummy, np.zeros((8, 8, 1, 1), np.float32), np.zeros(8, np.float32), 1, 1)
    p.finalize()
    p.write(sl, x)
    cal = Q.HistogramCalibrator()
    n = cal.bins_for(float(x.abs().max()))
    hist = torch.zeros(n, dtype=torch.int32, device="cuda")
    p.slice_histogram(sl, 3, 1.0 / cal.width, hist)
    torch.cuda.synchronize()
    ref = Q.HistogramCalibrator()
    ref.collect_host(x.permute(0, 2, 3, 1).numpy())
    assert np.array_equal(hist.cpu().numpy().astype(np.int64), ref.hist) and int(hist.sum()) == x.numel()

    m = uyd.UninaYoloB200.from_yaml().init_synthetic(seed=0).cuda()
    frames = oi.seeded_frames(6, 320, seed=17).cuda()
    a_max = m.calibrate_int8(frames, enable=False, method="max", batch_size=4)
    a_ent = m.calibrate_int8(frames, enable=True, method="histogram", batch_size=4)     # the reference's default
    assert a_ent.keys() == a_max.keys() and len(a_ent) == 158
    ratio = np.array([a_ent[k][0] / a_max[k][0] for k in a_ent])
    assert ratio.max() <= 1.0 + 2e-3 and np.median(ratio) < 1.0 and ratio.min() > 0.05
    xs = frames[:2].contiguous()
    _, raws = m(xs)
    pl = m.plan_for(xs)
    l2 = pl.read(pl.layer_outputs[2], 2).cpu()
    ref_m = yg.DetectionModel(yg.default_yaml_path())
    ref_m.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()}, strict=True)
    want = Int8Graph(ref_m.eval(), a_ent).forward_from({2: l2})
    for a, b in zip(raws, want):
        assert a.cpu().numpy().tobytes() == b.numpy().tobytes()
