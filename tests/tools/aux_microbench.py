#!/usr/bin/env python
"""CUDA-event timings of the evaluation / augmentation kernels outside the training step (region-based metrics,
thin-plate-spline warp) at the configs' sizes, with the oracle timed beside them on a bounded sample.

  python tests/tools/aux_microbench.py > gpurun_out/aux_microbench.txt
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dnncancerannotator_b200 import data_tail as DT                  # noqa: E402
from dnncancerannotator_b200 import native as N                      # noqa: E402
from dnncancerannotator_b200.synthetic import make_slices           # noqa: E402
from dnncancerannotator_b200.utils import metrics as M              # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def region(B, S, thresholds, resize):
    from scipy import ndimage
    _, y = make_slices(B, S, S, 3, seed=5)
    rng = np.random.default_rng(0)
    p = ndimage.gaussian_filter(rng.normal(size=y.shape), (0, 3, 3))
    p = (1 / (1 + np.exp(-(p * 12 + 3 * y - 1.5)))).astype(np.float32)
    yd, pd = torch.from_numpy(y).cuda(), torch.from_numpy(p).cuda()[..., None]
    eng = M.RegionCounts(thresholds, 0.3, resize, 5, 'cuda')
    with N.Profiler() as prof:
        eng.update(yd, pd)
    parts = {k: round(v['ms'], 4) for k, v in prof.summary().items()}
    ms = timed(lambda: eng.update(yd, pd))
    eng.check()
    T = len(thresholds)
    px = B * S * S
    print(f'region_confusion B={B} {S}x{S} T={T} resize={resize}: {ms:.3f} ms/batch = {B / ms * 1e3:,.0f} slices/s, '
          f'{px * (T + 1) / ms / 1e6:,.1f} Gpixel-planes/s; input bytes {px * 8 / 1e6:.1f} MB', flush=True)
    return y, p


def region_cpu(y, p, thresholds, resize, n):
    from oracle import ref_region as rr
    t0 = time.perf_counter()
    rr.get_tp_fn_fp(y[:n], p[:n, ..., None], thresholds, 0.3, resize)
    dt = time.perf_counter() - t0
    print(f'   oracle (numpy restatement, 1 thread): {n} slice(s) in {dt:.2f} s = {n / dt:.2f} slices/s', flush=True)


def warp(B, S, C, P, max_diff, stddev):
    rng = np.random.default_rng(1)
    img = torch.from_numpy(rng.uniform(size=(B, S, S, C)).astype(np.float32)).cuda()
    src, dst = DT.draw_control_points(rng, B, S, P, max_diff, stddev)
    src, dst = torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda()
    with N.Profiler() as prof:
        DT.sparse_image_warp(img, src, dst, return_flow=False)
    parts = {k: round(v['ms'], 4) for k, v in prof.summary().items()}
    ms = timed(lambda: DT.sparse_image_warp(img, src, dst, return_flow=False), reps=10)
    print(f'tps warp B={B} {S}x{S}x{C} P={P}: {ms:.3f} ms/batch = {B / ms * 1e3:,.0f} slices/s  (fit / warp kernels: {parts}); '
          f'{B * S * S * P / ms / 1e6:,.1f} G phi evaluations/s (FP64)', flush=True)
    return img, src, dst


def warp_cpu(img, src, dst, n):
    from oracle import ref_warp as rw
    a, s, d = img[:n].cpu().numpy(), src[:n].cpu().numpy(), dst[:n].cpu().numpy()
    for dt_, name in ((np.float32, 'float32 (the reference precision)'), (np.float64, 'float64')):
        t0 = time.perf_counter()
        rw.sparse_image_warp(a, s, d, dtype=dt_)
        dt = time.perf_counter() - t0
        print(f'   oracle {name}, numpy/BLAS on {torch.get_num_threads()} threads: {n} slice(s) in {dt:.2f} s = {n / dt:.2f} slices/s',
              flush=True)


if __name__ == '__main__':
    print(torch.cuda.get_device_name(0))
    y, p = region(32, 256, [0.8], 0.5)                  # configs/additionals/metrics.yaml:24-59
    region_cpu(y, p, [0.8], 0.5, 4)
    y, p = region(32, 256, [0.8], 1.0)
    region(64, 256, [i / 9 if i else 0.001 for i in range(10)], 1.0)   # Visualizer region PR curve, eval batch 64
    region(256, 256, [0.8], 1.0)
    img, src, dst = warp(10, 256, 6, 100, 5, 2.0)        # random_warp defaults, process_in_batch=10 (data.py:628,719)
    warp_cpu(img, src, dst, 2)
    warp(15, 256, 6, 150, 15, 20.0)                      # configs/additionals/augment_options.yaml
    warp(32, 256, 6, 100, 5, 2.0)
    warp(10, 512, 6, 100, 5, 2.0)
