import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dnncancerannotator_b200 import native as N
N.lib()
bf = torch.bfloat16
ws = torch.empty(64 << 20, dtype=torch.uint8, device='cuda')
def t(fn, reps=10):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
lib = N.lib()
for (B, H, ca, co) in [(32, 256, 1, 16), (32, 256, 16, 16), (32, 128, 16, 32), (32, 128, 32, 32), (32, 64, 64, 64)]:
    x = torch.randn(B, H, H, ca, device='cuda').to(bf)
    y = torch.empty(B, H, H, co, device='cuda', dtype=bf)
    w = torch.randn(3, 3, ca, co, device='cuda') * 0.1
    b = torch.randn(co, device='cuda')
    stats = torch.zeros(2 * co, dtype=torch.float64, device='cuda')
    xv, yv = N.tensor_view(x), N.tensor_view(y)
    for sp, tag in ((None, 'no stats'), (N.ptr(stats), 'stats')):
        lib.dnnca_debug_family_count(2, 1); lib.dnnca_debug_family_count(0, 1); lib.dnnca_debug_family_count(1, 1)
        us = t(lambda: N.call('dnnca_conv2d_fprop', None, C.byref(xv), None, N.ptr(w), N.ptr(b), C.byref(yv), 3, N.ACT_RELU, 0.0, sp, N.ptr(ws), ws.numel()))
        print(f'fprop [{ca}->{co}@{H}] B={B} {tag}: {us:8.1f} us   families tc={lib.dnnca_debug_family_count(2,0)} small={lib.dnnca_debug_family_count(1,0)} generic={lib.dnnca_debug_family_count(0,0)}', flush=True)
    us = t(lambda: N.call('dnnca_channel_stats', None, C.byref(yv), N.ptr(stats)))
    print(f'   channel_stats [{co}@{H}]: {us:8.1f} us')
