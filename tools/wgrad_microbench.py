#!/usr/bin/env python
"""CUDA-event time of dnnca_conv2d_wgrad at the layer shapes of configs/unet_big.yaml (batch 16), with and without
the bias gradient (a separate channel-sum pass).  DNNCA_DISABLE_WGRAD_HALO=1 selects the first-generation kernel."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dnncancerannotator_b200 import native as N
N.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
SHAPES = [(256, 64, 0, 64), (256, 64, 64, 64), (128, 64, 0, 128), (128, 128, 0, 128), (128, 128, 128, 128), (64, 256, 0, 256),
          (64, 256, 256, 256), (32, 512, 0, 512), (32, 512, 512, 512)]
bf = torch.bfloat16
for H, ca, cb, co in SHAPES:
    xa = torch.randn(B, H, H, ca, device='cuda').to(bf)
    xb = torch.randn(B, H, H, cb, device='cuda').to(bf) if cb else None
    dz = torch.randn(B, H, H, co, device='cuda').to(bf)
    dw = torch.zeros(3, 3, ca + cb, co, device='cuda')
    db = torch.zeros(co, device='cuda')
    xav, dzv = N.tensor_view(xa), N.tensor_view(dz)
    xbp = C.byref(N.tensor_view(xb)) if cb else None
    for with_db in (False, True):
        fn = lambda: N.call('dnnca_conv2d_wgrad', None, C.byref(xav), xbp, C.byref(dzv), N.ptr(dw), N.ptr(db) if with_db else None, 3)
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 2.0 * B * H * H * (ca + cb) * co * 9
        print(f'wgrad [{ca}+{cb}->{co}@{H}] db={int(with_db)} {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TFLOP/s', flush=True)
