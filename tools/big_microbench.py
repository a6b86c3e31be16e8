#!/usr/bin/env python
"""Per-layer microbenchmark of the 3x3 conv C-ABI calls at the shapes of configs/unet_big.yaml (BatchNorm statistics
fused as in the model): CUDA-event time, algorithmic GB/s and TFLOP/s of fprop / dgrad / wgrad, product path only.

  python tools/big_microbench.py [--batch 32] [--reps 10] [--only fprop|dgrad|wgrad] [--layers 1,16]
"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dnncancerannotator_b200 import native as N  # noqa: E402

# (H, Cx, Cx2, Cout) of every 3x3 conv of configs/unet_big.yaml
LAYERS = [(256, 3, 0, 64), (256, 64, 0, 64), (128, 64, 0, 128), (128, 128, 0, 128), (64, 128, 0, 256), (64, 256, 0, 256),
          (32, 256, 0, 512), (32, 512, 0, 512), (16, 512, 0, 1024), (16, 1024, 0, 1024), (32, 512, 512, 512), (32, 512, 0, 512),
          (64, 256, 256, 256), (64, 256, 0, 256), (128, 128, 128, 128), (128, 128, 0, 128), (256, 64, 64, 64), (256, 64, 0, 64)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--only', default='')
    ap.add_argument('--layers', default='')
    ap.add_argument('--no-stats', action='store_true')
    ap.add_argument('--mask', action='store_true', help='dgrad applies a ReLU mask (models without BatchNorm)')
    args = ap.parse_args()
    N.lib()
    B = args.batch
    sel = [int(v) for v in args.layers.split(',')] if args.layers else range(len(LAYERS))
    bf = torch.bfloat16
    ws = torch.empty(64 << 20, dtype=torch.uint8, device='cuda')
    tot = {}
    for li in sel:
        H, ca, cb, co = LAYERS[li]
        cin = ca + cb
        cpad = 8 if ca < 8 else ca
        xa_buf = torch.randn(B, H, H, cpad, device='cuda').to(bf)
        xb = torch.randn(B, H, H, cb, device='cuda').to(bf) if cb else None
        y = torch.empty(B, H, H, co, device='cuda', dtype=bf)
        dz = torch.randn(B, H, H, co, device='cuda').to(bf)
        dxa = torch.empty_like(xa_buf)
        dxb = torch.empty_like(xb) if cb else None
        w = torch.randn(3, 3, cin, co, device='cuda') * 0.05
        b = torch.randn(co, device='cuda')
        dw = torch.zeros_like(w)
        db = torch.zeros_like(b)
        stats = torch.zeros(2 * co, dtype=torch.float64, device='cuda')
        xav, dxav = N.tensor_view(xa_buf, 0, ca), N.tensor_view(dxa, 0, ca)
        yv, dzv = N.tensor_view(y), N.tensor_view(dz)
        xbp = C.byref(N.tensor_view(xb)) if cb else None
        dxbp = C.byref(N.tensor_view(dxb)) if cb else None
        px = B * H * H
        flops = 2.0 * px * cin * 9 * co
        sp = None if args.no_stats else N.ptr(stats)
        calls = {
            'fprop': (lambda: N.call('dnnca_conv2d_fprop', None, C.byref(xav), xbp, N.ptr(w), N.ptr(b), C.byref(yv), 3,
                                     N.ACT_NONE, 0.0, sp, N.ptr(ws), ws.numel()), px * (cin + co) * 2),
            'dgrad': (lambda: N.call('dnnca_conv2d_dgrad', None, C.byref(dzv), N.ptr(w), C.byref(dxav), dxbp, 3,
                                     C.byref(xav) if args.mask else None, N.ACT_RELU if args.mask else N.ACT_NONE, 0.0,
                                     N.ptr(ws), ws.numel()), px * (co + cin + (ca if args.mask else 0)) * 2),
            'wgrad': (lambda: N.call('dnnca_conv2d_wgrad', None, C.byref(xav), xbp, C.byref(dzv), N.ptr(dw), N.ptr(db), 3),
                      px * (cin + co) * 2),
        }
        for name, (fn, nbytes) in calls.items():
            if args.only and name != args.only:
                continue
            if name == 'dgrad' and ca < 8:
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
            print(f'L{li:02d} {name} [{ca}+{cb}->{co}@{H}] {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s  {flops / ms / 1e9:7.1f} TFLOP/s',
                  flush=True)
    print('total ms:', {k: round(v, 3) for k, v in tot.items()})


if __name__ == '__main__':
    main()
