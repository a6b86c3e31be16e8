#!/usr/bin/env python
"""Microbenchmark of the ConvT 2x2/s2 C-ABI calls at the shapes of configs/unet_big.yaml: CUDA-event time, algorithmic
GB/s and TFLOP/s of fprop / dgrad / wgrad, product path only.

  python tools/tconv_microbench.py [--batch 32] [--reps 10] [--only fprop|dgrad|wgrad] [--layers 0,3]
"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dnncancerannotator_b200 import native as N  # noqa: E402

# (input H, Cin, Cout) of the four up-sampling layers of configs/unet_big.yaml
LAYERS = [(128, 128, 64), (64, 256, 128), (32, 512, 256), (16, 1024, 512)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--only', default='')
    ap.add_argument('--layers', default='')
    args = ap.parse_args()
    N.lib()
    B = args.batch
    sel = [int(v) for v in args.layers.split(',')] if args.layers else range(len(LAYERS))
    bf = torch.bfloat16
    ws = torch.empty(64 << 20, dtype=torch.uint8, device='cuda')
    tot = {}
    for li in sel:
        H, cin, co = LAYERS[li]
        x = torch.randn(B, H, H, cin, device='cuda').to(bf)
        y = torch.empty(B, 2 * H, 2 * H, co, device='cuda', dtype=bf)
        dy = torch.randn(B, 2 * H, 2 * H, co, device='cuda').to(bf)
        dx = torch.empty_like(x)
        k = torch.randn(2, 2, co, cin, device='cuda') * 0.1
        b = torch.randn(co, device='cuda')
        dk = torch.zeros_like(k)
        db = torch.zeros_like(b)
        xv, yv, dyv, dxv = (N.tensor_view(t) for t in (x, y, dy, dx))
        px = B * H * H
        flops = 2.0 * px * cin * 4 * co
        calls = {
            'fprop': (lambda: N.call('dnnca_convtranspose2x2_fprop', None, C.byref(xv), N.ptr(k), N.ptr(b), C.byref(yv), None,
                                     N.ptr(ws), ws.numel()), px * (cin + 4 * co) * 2),
            'dgrad': (lambda: N.call('dnnca_convtranspose2x2_dgrad', None, C.byref(dyv), N.ptr(k), C.byref(dxv), C.byref(xv),
                                     N.ACT_RELU, 0.0, N.ptr(ws), ws.numel()), px * (4 * co + 2 * cin) * 2),
            'wgrad': (lambda: N.call('dnnca_convtranspose2x2_wgrad', None, C.byref(xv), C.byref(dyv), N.ptr(dk), N.ptr(db)),
                      px * (cin + 4 * co) * 2),
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
            print(f'T{li} {name} [{cin}->{co}@{H}] {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s  {flops / ms / 1e9:7.1f} TFLOP/s',
                  flush=True)
    print('total ms:', {k: round(v, 3) for k, v in tot.items()})


if __name__ == '__main__':
    main()
