"""Evaluation consumer: the CUDA matcher vs the restated reference loops."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene(B, seed, n_gt=40, max_det=300):
    """Ground truths with many small boxes, detections = jittered copies + clutter, in confidence order."""
    g = torch.Generator().manual_seed(seed)
    gts, dets = [], []
    for b in range(B):
        k = int(torch.randint(0, n_gt, (1,), generator=g)) if b else n_gt
        c = torch.rand(k, 2, generator=g) * 600 + 20
        wh = torch.where(torch.rand(k, 1, generator=g) < 0.6, torch.rand(k, 2, generator=g) * 10 + 4, torch.rand(k, 2, generator=g) * 60 + 16)
        gt = torch.cat((torch.randint(0, 4, (k, 1), generator=g).float(), c - wh / 2, c + wh / 2), 1)
        keep = torch.rand(k, generator=g) < 0.8
        jit = gt[keep, 1:] + torch.randn(int(keep.sum()), 4, generator=g) * 1.2
        cls = torch.where(torch.rand(int(keep.sum()), generator=g) < 0.9, gt[keep, 0], torch.randint(0, 4, (int(keep.sum()),), generator=g).float())
        m = int(torch.randint(0, 60, (1,), generator=g))
        cc = torch.rand(m, 2, generator=g) * 600 + 20
        cw = torch.rand(m, 2, generator=g) * 20 + 3
        clutter = torch.cat((cc - cw / 2, cc + cw / 2), 1)
        boxes = torch.cat((jit, clutter, jit[: len(jit) // 3] + 0.5))          # duplicates compete for the same ground truth
        cl = torch.cat((cls, torch.randint(0, 4, (m,), generator=g).float(), cls[: len(jit) // 3]))
        conf = torch.rand(len(boxes), generator=g)
        order = torch.argsort(conf, descending=True, stable=True)
        d = torch.cat((boxes, conf[:, None], cl[:, None]), 1)[order][:max_det]
        if b == 2:
            d = d[:0]                                                        # an image without detections
        if b == 3:
            gt = gt[:0]                                                      # an image without labels
        gts.append(gt)
        dets.append(d)
    return dets, gts


def test_eval_consumer_matches_reference_loops():
    from unina_yolo_dla_b200.evaluate import DetectionEvaluator
    from oracle import evalref as er

    B, max_det, gmax = 6, 300, 48
    ev = DetectionEvaluator(size_threshold=15, small_iou=0.45, match_iou=0.5)
    all_scores, tot = [], np.zeros(3, np.int64)
    for seed in (1, 2):
        dets, gts = _scene(B, seed)
        det = torch.zeros(B, max_det, 6)
        cnt = torch.zeros(B, dtype=torch.int32)
        gt = torch.zeros(B, gmax, 5)
        gcnt = torch.zeros(B, dtype=torch.int32)
        for b in range(B):
            det[b, : len(dets[b])] = dets[b]
            cnt[b] = len(dets[b])
            gt[b, : len(gts[b])] = gts[b]
            gcnt[b] = len(gts[b])
        ev.update(det.cuda(), cnt.cuda(), gt.cuda(), gcnt.cuda())
        tot += np.asarray(er.small_object_counts(dets, gts, 15, 0.45))
        all_scores.append(er.conformal_scores(dets, gts, 0.5))
    m = ev.small_object_metrics()
    assert (m["tp"], m["fp"], m["fn"]) == tuple(int(v) for v in tot) and m["tp"] > 20 and m["fp"] > 5 and m["fn"] > 5
    want = np.concatenate(all_scores)
    got = ev.nonconformity_scores().cpu().numpy()
    assert got.shape == want.shape and len(want) > 50
    assert np.array_equal(got, want)                                   # same pairs, same fp32 IoU
    cp = ev.conformal(alpha=0.10)
    assert abs(cp["q_hat"] - er.conformal_quantile(want, 0.10)) < 1e-9 and cp["num_calibration_samples"] == len(want)
    p = m["tp"] / (m["tp"] + m["fp"] + 1e-7)
    assert m["metrics/small_precision"] == p
