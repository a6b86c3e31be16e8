"""Seeded synthetic slices with the reference's input contract.

The reference feeds ``x`` float32 ``[B,H,W,C]`` in {k/255} (uint8 / 255,
``annotator/data.py:193-206``) and ``y`` float32 ``[B,H,W]`` in [0,1]
(label channel / 255, ``data.py:766-788``; all-zero for healthy exams,
``data.py:417-421``).  Labels are unions of filled discs like the region-metric
tests draw (``annotator/tests/test_region_metrics.py:318-336``); every 4th slice
is all-zero so that the ``positive_rate == 0`` branch (``losses.py:27``) and the
low positive rate (1-3 %) of real data are exercised.
"""
import numpy as np


def make_slices(batch, height=256, width=256, channels=3, seed=1234, as_uint8=False):
    """Returns (x, y): float32 ``[B,H,W,C]`` / ``[B,H,W]`` (or the uint8 pre-/255 form)."""
    rng = np.random.default_rng(seed)
    x8 = rng.integers(0, 256, (batch, height, width, channels), dtype=np.uint8)
    y8 = np.zeros((batch, height, width), np.uint8)
    yy, xx = np.mgrid[0:height, 0:width]
    rmax = max(2, min(height, width) // 10)
    rmin = max(1, rmax // 5)
    for b in range(batch):
        if b % 4 == 3:
            continue  # healthy slice
        for _ in range(int(rng.integers(1, 4))):
            r = rng.uniform(rmin, rmax)
            cy, cx = rng.uniform(0, height), rng.uniform(0, width)
            y8[b][(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 255
    if as_uint8:
        return x8, y8
    return x8.astype(np.float32) / np.float32(255.0), y8.astype(np.float32) / np.float32(255.0)
