"""Numpy-only test helpers (usable in CPU worker processes)."""
import numpy as np


def synth_predictions(B, nc, A, seed=0, frac_conf=0.2, img=640.0, cluster=False):
    """Synthetic y[B,4+nc,A] with overlapping boxes and a controllable share above conf."""
    rng = np.random.default_rng(seed)
    y = np.zeros((B, 4 + nc, A), np.float32)
    if cluster:
        centers = rng.uniform(40, img - 40, (B, 2, 64))
        pick = rng.integers(0, 64, (B, A))
        cxy = np.take_along_axis(centers, pick[:, None, :].repeat(2, 1), 2) + rng.normal(0, 6, (B, 2, A))
    else:
        cxy = rng.uniform(0, img, (B, 2, A))
    y[:, 0:2] = cxy
    y[:, 2:4] = rng.uniform(8, 96, (B, 2, A))
    sc = rng.uniform(0, 1, (B, nc, A)).astype(np.float32)
    lift = rng.uniform(0, 1, (B, 1, A)) < frac_conf
    y[:, 4:] = np.where(lift, 0.25 + 0.75 * sc, 0.2 * sc)
    return y
