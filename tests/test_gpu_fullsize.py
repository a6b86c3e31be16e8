"""The conv kernels at BASELINE.json's FULL sizes (per-GPU batch 256 of 256x256 for unet.yaml, batch 32 for the wide
configs), where the oracle cannot run a whole tensor in seconds: size-independent properties plus oracle spot checks.

For every layer shape, through the C ABI, bf16:

* **oracle on a sample**: the first, middle and last image of the batch against the torch-CPU oracle conv of the same
  (bf16-valued) inputs -- the launch geometry is the full-size one, the checked values include every image border;
* **batch-split invariance**: the second half of the batch computed by a separate call equals the same images inside
  the full-batch call BIT FOR BIT (an output element's accumulation order does not depend on the tiling of the batch);
* **adjoint identities** (fprop, dgrad and wgrad are the three faces of one bilinear form): with dz := y,
  <y, y> = <x, dgrad(y, W)> = <W, wgrad(x, y)> up to bf16 rounding of the stored tensors / weight operands;
* **checksums**: the BatchNorm statistics the fprop epilogue emits (sum, sum of squares per channel) and the bias
  gradient of wgrad against fp64 reductions of the stored tensors.

Tolerances (floating point): 1.2e-2 of max|ref| per element for the sample (the per-op bf16 bound of test_gpu_ops.py);
5e-3 relative for the adjoint sums (fprop of the few-channel row kernels uses hi|lo = fp32-accurate weights, dgrad a
single bf16 band: the two sides differ by the weight rounding, ~2^-9 / sqrt(K) per term); 1e-4 for the checksums.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import ref_ops as ops

pytestmark = pytest.mark.gpu

# (tag, n, h, w, c_x, c_x2, cout)
SHAPES = [
    ('unet.yaml d0.conv1 / u2.conv1', 256, 256, 256, 3, 0, 3),
    ('unet.yaml u2.conv0 (virtual concat)', 256, 256, 256, 3, 3, 3),
    ('unet.yaml d1.conv1', 256, 128, 128, 6, 0, 6),
    ('unet.yaml u0.conv0 (virtual concat)', 256, 64, 64, 12, 12, 12),
    ('mulmo_unet encoder conv1 @256', 32, 256, 256, 16, 0, 16),
    ('mulmo_unet encoder conv1 @128', 32, 128, 128, 32, 0, 32),
    ('unet_big d0.conv1', 32, 256, 256, 64, 0, 64),
    ('unet_big u3.conv0 (virtual concat)', 32, 256, 256, 64, 64, 64),
    ('unet_big d1.conv1', 32, 128, 128, 128, 0, 128),
    ('unet_big d3.conv1', 32, 32, 32, 512, 0, 512),
]


@pytest.fixture(scope='module')
def N():
    from dnncancerannotator_b200 import native
    native.lib()
    return native


def _dot(a, b):
    return float((a.double() * b.double()).sum())


@pytest.mark.parametrize('shape', SHAPES, ids=[s[0] for s in SHAPES])
def test_conv_full_size_properties(N, shape):
    tag, n, h, w, ca, cb, cout = shape
    cin, k = ca + cb, 3
    lib = N.lib()
    g = torch.Generator(device='cuda').manual_seed(n * 7 + h + cin)
    xa = torch.randn(n, h, w, ca, generator=g, device='cuda').bfloat16()
    xb = torch.randn(n, h, w, cb, generator=g, device='cuda').bfloat16() if cb else None
    wt = (torch.randn(k, k, cin, cout, generator=g, device='cuda') / np.sqrt(k * k * cin)).float().contiguous()
    bias = torch.zeros(cout, device='cuda')
    ws = torch.zeros(int(lib.dnnca_conv_workspace_bytes(k * k, cin, cout)) + 16, dtype=torch.uint8, device='cuda')
    WS = (N.ptr(ws), ws.numel())

    def fprop(xa_, xb_, y_, stats_=None):
        xav, yv = N.tensor_view(xa_), N.tensor_view(y_)
        xbv = N.tensor_view(xb_) if xb_ is not None else None
        N.call('dnnca_conv2d_fprop', N.stream_ptr(), C.byref(xav), C.byref(xbv) if xbv is not None else None, N.ptr(wt), N.ptr(bias),
               C.byref(yv), k, N.ACT_NONE, 0.0, N.ptr(stats_) if stats_ is not None else None, *WS)

    lib.dnnca_debug_family_count(0, 1)
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device='cuda')
    stats = torch.zeros(2 * cout, dtype=torch.float64, device='cuda')
    fprop(xa, xb, y, stats)
    torch.cuda.synchronize()
    assert int(lib.dnnca_debug_family_count(0, 0)) == 0, 'a full-size headline shape fell back to the generic CUDA-core kernel'

    # ---- oracle on a sample of the batch
    x_all = xa if xb is None else torch.cat([xa, xb], -1)
    for i in (0, n // 2, n - 1):
        ref = ops.conv2d(x_all[i:i + 1].float().cpu(), wt.cpu(), bias.cpu()).numpy()
        got = y[i:i + 1].float().cpu().numpy()
        err = np.abs(got - ref).max()
        assert err <= 1.2e-2 * np.abs(ref).max(), (tag, i, err, np.abs(ref).max())

    # ---- batch-split invariance, bit for bit
    y2 = torch.empty(n - n // 2, h, w, cout, dtype=torch.bfloat16, device='cuda')
    fprop(xa[n // 2:], xb[n // 2:] if xb is not None else None, y2)
    torch.cuda.synchronize()
    assert torch.equal(y2, y[n // 2:]), (tag, 'fprop of a half batch differs from the same images in the full batch')

    # ---- checksums of the epilogue statistics
    yd = y.double()
    s1, s2 = yd.sum((0, 1, 2)), (yd * yd).sum((0, 1, 2))
    assert torch.allclose(stats[:cout], s1, rtol=1e-4, atol=1e-4 * float(s2.max().sqrt()) * np.sqrt(n * h * w)), tag
    assert torch.allclose(stats[cout:], s2, rtol=1e-4), tag

    # ---- adjoint identities with dz := y
    dz = y
    dxa = torch.empty_like(xa)
    dxb = torch.empty_like(xb) if xb is not None else None
    dzv, dxav = N.tensor_view(dz), N.tensor_view(dxa)
    dxbv = N.tensor_view(dxb) if dxb is not None else None
    N.call('dnnca_conv2d_dgrad', N.stream_ptr(), C.byref(dzv), N.ptr(wt), C.byref(dxav), C.byref(dxbv) if dxbv is not None else None, k,
           None, N.ACT_NONE, 0.0, *WS)
    dw = torch.zeros(k, k, cin, cout, dtype=torch.float32, device='cuda')
    db = torch.zeros(cout, dtype=torch.float32, device='cuda')
    xav = N.tensor_view(xa)
    xbv = N.tensor_view(xb) if xb is not None else None
    N.call('dnnca_conv2d_wgrad', N.stream_ptr(), C.byref(xav), C.byref(xbv) if xbv is not None else None, C.byref(dzv), N.ptr(dw),
           N.ptr(db), k)
    torch.cuda.synchronize()
    assert int(lib.dnnca_debug_family_count(0, 0)) == 0, 'dgrad / wgrad of a full-size headline shape fell back to the generic kernel'
    yy = float((yd * yd).sum())
    via_dgrad = _dot(xa, dxa) + (_dot(xb, dxb) if xb is not None else 0.0)
    via_wgrad = _dot(wt, dw)
    assert abs(via_dgrad - yy) <= 5e-3 * yy, (tag, yy, via_dgrad)
    assert abs(via_wgrad - yy) <= 5e-3 * yy, (tag, yy, via_wgrad)
    # bias gradient = per-channel sum of dz
    assert torch.allclose(db.double(), s1, rtol=1e-4, atol=1e-4 * float(s2.max().sqrt()) * np.sqrt(n * h * w)), tag

    # ---- dgrad is batch-split invariant too
    dx2 = torch.empty(n - n // 2, h, w, ca, dtype=torch.bfloat16, device='cuda')
    dzh, dx2v = N.tensor_view(dz[n // 2:]), N.tensor_view(dx2)
    dxb2 = torch.empty(n - n // 2, h, w, cb, dtype=torch.bfloat16, device='cuda') if cb else None
    dxb2v = N.tensor_view(dxb2) if cb else None
    N.call('dnnca_conv2d_dgrad', N.stream_ptr(), C.byref(dzh), N.ptr(wt), C.byref(dx2v), C.byref(dxb2v) if cb else None, k, None,
           N.ACT_NONE, 0.0, *WS)
    torch.cuda.synchronize()
    assert torch.equal(dx2, dxa[n // 2:]), (tag, 'dgrad of a half batch differs from the same images in the full batch')
    if cb:
        assert torch.equal(dxb2, dxb[n // 2:]), tag


@pytest.mark.parametrize('n,h,c', [(256, 256, 3), (256, 128, 6), (32, 256, 64), (32, 128, 128)])
def test_maxpool_full_size_bit_exact(N, n, h, c):
    """MaxPool2D([2,2], 2) at the configs' sizes: values bit-exact against torch's max_pool2d of the same bf16 tensor;
    every index points at an element equal to the maximum."""
    g = torch.Generator(device='cuda').manual_seed(n + h + c)
    x = torch.randn(n, h, h, c, generator=g, device='cuda').bfloat16()
    y = torch.empty(n, h // 2, h // 2, c, dtype=torch.bfloat16, device='cuda')
    idx = torch.empty(n, h // 2, h // 2, c, dtype=torch.uint8, device='cuda')
    xv, yv = N.tensor_view(x), N.tensor_view(y)
    N.call('dnnca_maxpool2x2_fwd', N.stream_ptr(), C.byref(xv), C.byref(yv), N.ptr(idx), None)
    torch.cuda.synchronize()
    ref = torch.nn.functional.max_pool2d(x.permute(0, 3, 1, 2).float(), 2, 2).permute(0, 2, 3, 1).bfloat16()
    assert torch.equal(y, ref)
    assert int(idx.max()) <= 3
    # the indexed element is the maximum (first-maximum tie-breaking is pinned by the small bit-exact tests)
    win = x.view(n, h // 2, 2, h // 2, 2, c).permute(0, 1, 3, 5, 2, 4).reshape(n, h // 2, h // 2, c, 4)
    picked = torch.gather(win, 4, idx.long().unsqueeze(-1)).squeeze(-1)
    assert torch.equal(picked, y)


def test_head_loss_full_size_against_fp64_formula(N):
    """head 1x1 + sigmoid + weighted BCE (losses.py:17-37) on a full unet.yaml batch (256 x 256 x 256 x 3 features):
    per-sample losses and the summed gradients against the formula evaluated in fp64 with torch on the device."""
    from dnncancerannotator_b200.synthetic import make_slices
    n, h, F = 256, 256, 3
    g = torch.Generator(device='cuda').manual_seed(5)
    f = torch.randn(n, h, h, F, generator=g, device='cuda').bfloat16()
    _, ynp = make_slices(n, h, h, 3, seed=9)
    y = torch.from_numpy(ynp).cuda()
    w = torch.tensor([0.7, -0.4, 0.2], device='cuda')
    b = torch.tensor([-1.5], device='cuda')
    stats = torch.zeros(16, dtype=torch.uint8, device='cuda')
    N.call('dnnca_label_stats_init', N.stream_ptr(), N.ptr(stats))
    N.call('dnnca_label_stats', N.stream_ptr(), N.ptr(y), y.numel(), N.ptr(stats))
    cfg = N.LossConfig(0.0, 0, 0.0, 3.0, 1.0 / (n * h * h))
    logits = torch.empty(n, h, h, 1, device='cuda')
    probs = torch.empty(n, h, h, 1, device='cuda')
    per = torch.zeros(n, device='cuda')
    df = torch.empty_like(f)
    dw = torch.zeros(F, device='cuda')
    db = torch.zeros(1, device='cuda')
    fv, dfv = N.tensor_view(f), N.tensor_view(df)
    N.call('dnnca_head_bce_fwd_bwd', N.stream_ptr(), C.byref(fv), N.ptr(w), N.ptr(b), N.ptr(y), N.ptr(stats), C.byref(cfg),
           N.ptr(logits), N.ptr(probs), N.ptr(per), C.byref(dfv), N.ACT_NONE, 0.0, N.ptr(dw), N.ptr(db))
    torch.cuda.synchronize()
    fd, yd = f.double(), y.double()
    z = (fd * w.double()).sum(-1) + b.double()
    r = yd.mean()
    weight = 3.0 * (1.0 / r if r > 0 else 1.0)
    mask = yd * (weight - 1.0) + 1.0
    loss = mask * (torch.clamp(z, min=0) - z * yd + torch.log1p(torch.exp(-z.abs())))
    ref_per = loss.mean((1, 2))
    assert torch.allclose(per.double(), ref_per, rtol=2e-5), float((per.double() - ref_per).abs().max())
    dzr = mask * (torch.sigmoid(z) - yd) / (n * h * h)
    assert abs(float(db) - float(dzr.sum())) <= 1e-4 * float(dzr.abs().sum())
    ref_dw = (dzr.unsqueeze(-1) * fd).sum((0, 1, 2))
    assert torch.allclose(dw.double(), ref_dw, rtol=1e-3, atol=1e-4 * float(ref_dw.abs().max()))
    assert torch.allclose(probs[..., 0].double(), torch.sigmoid(z), atol=1e-6)


@pytest.mark.parametrize('n,h,c', [(32, 256, 64), (32, 128, 128), (32, 256, 16), (32, 32, 512)])
def test_batchnorm_full_size_properties(N, n, h, c):
    """Training-mode BatchNormalization at the wide configs' sizes.  Forward: the normalised tensor has mean beta and
    variance gamma^2 * var / (var + eps) per channel (checked on the stored bf16 output against fp64 reductions of the
    stored input).  Backward without an activation mask: sum(dx) = 0 and sum(dx * xhat) = 0 per channel (the two
    directions BatchNorm projects out), dbeta = sum(dy), dgamma = sum(dy * xhat), and dx equals the bf16 rounding of the fp64
    closed form over the whole tensor."""
    g = torch.Generator(device='cuda').manual_seed(n + h + c)
    x = (torch.randn(n, h, h, c, generator=g, device='cuda') * 1.7 + 0.4).bfloat16()
    gamma = (torch.rand(c, generator=g, device='cuda') + 0.5).float()
    beta = (torch.randn(c, generator=g, device='cuda') * 0.1).float()
    count = n * h * h
    xv = N.tensor_view(x)
    stats = torch.zeros(4 * c, dtype=torch.float64, device='cuda')
    N.call('dnnca_channel_stats', N.stream_ptr(), C.byref(xv), N.ptr(stats))
    ss, mi = torch.zeros(2 * c, device='cuda'), torch.zeros(2 * c, device='cuda')
    mm, mv = torch.zeros(c, device='cuda'), torch.ones(c, device='cuda')
    N.call('dnnca_bn_finalize', N.stream_ptr(), N.ptr(stats), count, c, N.ptr(gamma), N.ptr(beta), 0.99, 1e-3, N.ptr(mm), N.ptr(mv),
           N.ptr(ss), N.ptr(mi))
    y = torch.empty_like(x)
    yv = N.tensor_view(y)
    N.call('dnnca_bn_apply', N.stream_ptr(), C.byref(xv), N.ptr(ss), C.byref(yv))
    torch.cuda.synchronize()
    xd = x.double()
    mean, var = xd.mean((0, 1, 2)), xd.var((0, 1, 2), unbiased=False)
    assert torch.allclose(stats[:c] / count, mean, rtol=1e-6, atol=1e-7)
    assert torch.allclose(mi[:c].double(), mean, rtol=1e-5, atol=1e-6) and torch.allclose(mi[c:].double(), (var + 1e-3).rsqrt(), rtol=1e-5)
    # moving statistics: 0.99 * old + 0.01 * batch, the variance unbiased ([TF-semantics])
    assert torch.allclose(mm.double(), 0.01 * mean, rtol=1e-4, atol=1e-7)
    assert torch.allclose(mv.double(), 0.99 + 0.01 * var * count / (count - 1), rtol=1e-5)
    yd = y.double()
    # the stored output is bf16 (ulp 2^-8 near 1): rounding averages out to ~1e-4 over a million pixels, not to zero
    assert torch.allclose(yd.mean((0, 1, 2)), beta.double(), atol=1e-3)
    assert torch.allclose(yd.var((0, 1, 2), unbiased=False), gamma.double() ** 2 * var / (var + 1e-3), rtol=5e-3)
    # backward
    dy = torch.randn(n, h, h, c, generator=g, device='cuda').bfloat16()
    dyv = N.tensor_view(dy)
    sums = stats[2 * c:]
    N.call('dnnca_bn_bwd_reduce', N.stream_ptr(), C.byref(xv), C.byref(dyv), N.ptr(mi), N.ptr(sums))
    dx = torch.empty_like(x)
    dxv = N.tensor_view(dx)
    dg, db = torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
    N.call('dnnca_bn_bwd_apply', N.stream_ptr(), C.byref(xv), C.byref(dyv), N.ptr(mi), N.ptr(gamma), N.ptr(sums), C.byref(dxv),
           N.ACT_NONE, 0.0, N.ptr(dg), N.ptr(db))
    torch.cuda.synchronize()
    xhat = (xd - mean) * (var + 1e-3).rsqrt()
    dyd, dxd = dy.double(), dx.double()
    assert torch.allclose(db.double(), dyd.sum((0, 1, 2)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(dg.double(), (dyd * xhat).sum((0, 1, 2)), rtol=1e-4, atol=1e-2)
    # dx against the closed form in fp64, whole tensor: the stored value is the bf16 rounding of it (measured: identical
    # in 99.999 % of the elements, never more than one bf16 ulp away)
    ref = (gamma.double() * (var + 1e-3).rsqrt()) * (dyd - dyd.mean((0, 1, 2)) - xhat * (dyd * xhat).mean((0, 1, 2)))
    assert bool(((dxd - ref).abs() <= 2.0 ** -7 * ref.abs() + 1e-6).all())
    assert float((dxd == ref.float().bfloat16().double()).double().mean()) >= 0.9999
    # the two directions BatchNorm projects out, up to the (slightly biased) bf16 rounding of a million stored values
    scale = dxd.abs().sum((0, 1, 2))
    assert bool((dxd.sum((0, 1, 2)).abs() <= 2e-3 * scale).all())
    assert bool(((dxd * xhat).sum((0, 1, 2)).abs() <= 2e-3 * scale).all())

@pytest.mark.parametrize('n,h,cin,cout', [(256, 128, 6, 3), (256, 32, 12, 12), (32, 128, 128, 64), (32, 16, 512, 512), (32, 128, 32, 16)])
def test_tconv_full_size_properties(N, n, h, cin, cout):
    """Conv2DTranspose 2x2 / 2 at the configs' sizes: oracle on sampled images, batch-split invariance bit for bit, and
    the adjoint identities <y, y> = <x, dgrad(y, K)> = <K, wgrad(x, y)> with a zero bias."""
    lib = N.lib()
    g = torch.Generator(device='cuda').manual_seed(n + h + cin)
    x = torch.randn(n, h, h, cin, generator=g, device='cuda').bfloat16()
    kt = (torch.randn(2, 2, cout, cin, generator=g, device='cuda') / np.sqrt(cin)).float().contiguous()
    bias = torch.zeros(cout, device='cuda')
    ws = torch.zeros(int(lib.dnnca_conv_workspace_bytes(4, cin, cout)) + 16, dtype=torch.uint8, device='cuda')
    WS = (N.ptr(ws), ws.numel())
    lib.dnnca_debug_family_count(0, 1)

    def fprop(x_, y_):
        xv, yv = N.tensor_view(x_), N.tensor_view(y_)
        N.call('dnnca_convtranspose2x2_fprop', N.stream_ptr(), C.byref(xv), N.ptr(kt), N.ptr(bias), C.byref(yv), None, *WS)
    y = torch.empty(n, 2 * h, 2 * h, cout, dtype=torch.bfloat16, device='cuda')
    fprop(x, y)
    torch.cuda.synchronize()
    for i in (0, n - 1):
        ref = ops.conv2d_transpose(x[i:i + 1].float().cpu(), kt.cpu(), bias.cpu()).numpy()
        got = y[i:i + 1].float().cpu().numpy()
        assert np.abs(got - ref).max() <= 1.2e-2 * np.abs(ref).max()
    y2 = torch.empty(n - n // 2, 2 * h, 2 * h, cout, dtype=torch.bfloat16, device='cuda')
    fprop(x[n // 2:], y2)
    torch.cuda.synchronize()
    assert torch.equal(y2, y[n // 2:])
    dx = torch.empty_like(x)
    yv, dxv, xv = N.tensor_view(y), N.tensor_view(dx), N.tensor_view(x)
    N.call('dnnca_convtranspose2x2_dgrad', N.stream_ptr(), C.byref(yv), N.ptr(kt), C.byref(dxv), None, N.ACT_NONE, 0.0, *WS)
    dk = torch.zeros(2, 2, cout, cin, dtype=torch.float32, device='cuda')
    db = torch.zeros(cout, dtype=torch.float32, device='cuda')
    N.call('dnnca_convtranspose2x2_wgrad', N.stream_ptr(), C.byref(xv), C.byref(yv), N.ptr(dk), N.ptr(db))
    torch.cuda.synchronize()
    assert int(lib.dnnca_debug_family_count(0, 0)) == 0, 'a full-size ConvT shape fell back to the generic kernel'
    yy = float((y.double() ** 2).sum())
    assert abs(_dot(x, dx) - yy) <= 5e-3 * yy, (yy, _dot(x, dx))
    assert abs(_dot(kt, dk) - yy) <= 5e-3 * yy, (yy, _dot(kt, dk))
    s1 = y.double().sum((0, 1, 2))
    assert torch.allclose(db.double(), s1, rtol=1e-4, atol=1e-4 * float((y.double() ** 2).sum((0, 1, 2)).max().sqrt()) * np.sqrt(n * 4 * h * h))
