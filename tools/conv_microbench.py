#!/usr/bin/env python
"""Per-layer microbenchmark of the conv C-ABI calls at the shapes of configs/unet.yaml (per-GPU batch 256):
CUDA-event time and algorithmic GB/s of fprop / dgrad / wgrad for every conv layer, product path only.

  python tools/conv_microbench.py [--batch 256] [--reps 10] [--only fprop|dgrad|wgrad] [--layers 0,3]
Used under ncu for per-kernel profiles (profiles/README.md); inputs are larger than L2.
"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dnncancerannotator_b200 import native as N  # noqa: E402

# (H, Cx, Cx2, Cout) of every 3x3 conv of configs/unet.yaml (C = 3 modalities)
LAYERS = [(256, 3, 0, 3), (256, 3, 0, 3), (128, 3, 0, 6), (128, 6, 0, 6), (64, 6, 0, 12), (64, 12, 0, 12),
          (64, 12, 12, 12), (64, 12, 0, 12), (128, 6, 6, 6), (128, 6, 0, 6), (256, 3, 3, 3), (256, 3, 0, 3)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--only', default='')
    ap.add_argument('--layers', default='')
    args = ap.parse_args()
    N.lib()
    B = args.batch
    sel = [int(v) for v in args.layers.split(',')] if args.layers else range(len(LAYERS))
    bf = torch.bfloat16
    tot = {}
    for li in sel:
        H, ca, cb, co = LAYERS[li]
        cin = ca + cb
        xa = torch.randn(B, H, H, ca, device='cuda').to(bf)
        xb = torch.randn(B, H, H, cb, device='cuda').to(bf) if cb else None
        y = torch.empty(B, H, H, co, device='cuda', dtype=bf)
        dz = torch.randn(B, H, H, co, device='cuda').to(bf)
        dxa = torch.empty_like(xa)
        dxb = torch.empty_like(xb) if cb else None
        w = torch.randn(3, 3, cin, co, device='cuda') * 0.1
        b = torch.randn(co, device='cuda')
        dw = torch.zeros_like(w)
        db = torch.zeros_like(b)
        xav, yv, dzv, dxav = (N.tensor_view(t) for t in (xa, y, dz, dxa))
        xbp = C.byref(N.tensor_view(xb)) if cb else None
        dxbp = C.byref(N.tensor_view(dxb)) if cb else None
        px = B * H * H
        calls = {
            'fprop': (lambda: N.call('dnnca_conv2d_fprop', None, C.byref(xav), xbp, N.ptr(w), N.ptr(b), C.byref(yv), 3,
                                     N.ACT_RELU, 0.0, None, None, 0), px * (cin + co) * 2),
            'dgrad': (lambda: N.call('dnnca_conv2d_dgrad', None, C.byref(dzv), N.ptr(w), C.byref(dxav), dxbp, 3,
                                     C.byref(xav), N.ACT_RELU, 0.0, None, 0), px * (co + cin + ca) * 2),
            'wgrad': (lambda: N.call('dnnca_conv2d_wgrad', None, C.byref(xav), xbp, C.byref(dzv), N.ptr(dw), N.ptr(db), 3),
                      px * (cin + co) * 2),
        }
        for name, (fn, nbytes) in calls.items():
            if args.only and name != args.only:
                continue
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            tot[name] = tot.get(name, 0.0) + ms
            print(f'L{li:02d} {name} [{ca}+{cb}->{co}@{H}] {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s', flush=True)
    print('total ms:', {k: round(v, 3) for k, v in tot.items()})


if __name__ == '__main__':
    main()
