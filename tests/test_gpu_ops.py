"""Per-op parity of the CUDA kernels (called through the C ABI) against the oracle.

Every case runs in fp32 (tight tolerance: same arithmetic, different summation order) and
bf16 (inputs rounded to bf16 first, so the only differences are the kernel's bf16 output
rounding and fp32 accumulation order).  Tensors are embedded as channel slices of wider
buffers to exercise the concat-slice views.  Conv shapes cover both the specialised
small-channel kernels and the shape-generic kernels (also forced via the debug hook).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import ref_ops as ops
from oracle import ref_numpy as rn

pytestmark = pytest.mark.gpu

DT = {'fp32': torch.float32, 'bf16': torch.bfloat16}
TOL = {'fp32': dict(rtol=2e-5, atol=2e-5), 'bf16': dict(rtol=1.2e-2, atol=1.2e-2)}


@pytest.fixture(scope='module')
def N():
    from dnncancerannotator_b200 import native
    native.lib()
    return native


def sync():
    torch.cuda.synchronize()


def embed(a, dtype, pad_lo=0, pad_hi=0, fill=7.0):
    """numpy NHWC array -> (torch buffer on GPU with extra channels, coff, c)."""
    n, h, w, c = a.shape
    buf = torch.full((n, h, w, pad_lo + c + pad_hi), fill, dtype=dtype, device='cuda')
    buf[..., pad_lo:pad_lo + c] = torch.from_numpy(a).to('cuda').to(dtype)
    return buf, pad_lo, c


def view(N, buf, coff, c):
    return N.tensor_view(buf, coff, c)


def q(a, dtype):
    """quantise a numpy array like the device storage does."""
    return torch.from_numpy(a).to(dtype).to(torch.float32).numpy()


_KEEP = []


@pytest.fixture(autouse=True)
def _keep_alive():
    """The C ABI sees raw pointers only: every device tensor made by dev() must outlive the launch."""
    _KEEP.clear()
    yield
    torch.cuda.synchronize()
    _KEEP.clear()


def dev(a, dtype=torch.float32):
    t = torch.from_numpy(np.ascontiguousarray(a)).to('cuda').to(dtype)
    _KEEP.append(t)
    return t


def close(got, ref, mode, scale=None):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    tol = TOL[mode]
    s = scale if scale is not None else max(np.abs(ref).max(), 1e-6)
    err = np.abs(got - ref).max()
    assert err <= tol['atol'] * s + 1e-12, f'max abs err {err:.3e} vs scale {s:.3e} (tol {tol["atol"]})'


CONV_SHAPES = [  # (n, h, w, c_x, c_x2, cout, k)
    (2, 16, 16, 3, 0, 3, 3), (1, 40, 72, 3, 0, 6, 3), (2, 8, 136, 6, 0, 6, 3), (1, 16, 16, 12, 12, 12, 3),
    (1, 12, 24, 12, 0, 12, 3), (2, 16, 16, 1, 0, 16, 3), (1, 16, 16, 5, 0, 3, 3), (1, 20, 264, 3, 3, 3, 3),
    (2, 24, 48, 6, 6, 6, 3), (1, 36, 40, 6, 0, 12, 3), (1, 16, 32, 8, 8, 8, 3), (2, 16, 16, 4, 0, 8, 3),
    (1, 10, 14, 7, 0, 9, 3), (1, 8, 8, 20, 0, 33, 3), (2, 9, 11, 17, 0, 8, 1), (1, 8, 8, 64, 0, 64, 3),
    (1, 10, 12, 7, 5, 9, 3),
]


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('variant', ['dense', 'sliced', 'generic'])
@pytest.mark.parametrize('shape', CONV_SHAPES)
def test_conv_fprop_dgrad_wgrad(N, mode, variant, shape):
    """dense: tensors are whole buffers (the TMA-staged small-channel kernels take the shapes they cover);
    sliced: every tensor is a channel slice of a wider buffer; generic: shape-generic kernels forced."""
    lib = N.lib()
    lib.dnnca_debug_family_count(1, 1)
    lib.dnnca_debug_family_count(2, 1)
    _conv_case(N, mode, variant, shape)
    n, h, w, ca, cb, cout, k = shape
    small = {(3, 0, 3), (3, 0, 6), (6, 0, 6), (12, 12, 12), (12, 0, 12), (1, 0, 16), (5, 0, 3), (3, 3, 3), (6, 6, 6),
             (6, 0, 12), (8, 8, 8), (4, 0, 8)}
    if variant == 'dense' and (ca, cb, cout) in small:
        es = 4 if mode == 'fp32' else 2
        assert all((w * c * es) % 16 == 0 for c in (ca, cb, cout) if c)
        # 3 fprop + wgrad + 2 dgrad, except first-layer style shapes whose dgrad is not instantiated
        taken = lib.dnnca_debug_family_count(1, 0) + lib.dnnca_debug_family_count(2, 0)
        assert taken >= 4, 'neither the row-Toeplitz tcgen05 kernels nor the TMA/FFMA2 small-channel kernels took this shape'


def _conv_case(N, mode, variant, shape):
    n, h, w, ca, cb, cout, k = shape
    cin = ca + cb
    dt = DT[mode]
    rng = np.random.default_rng(hash(shape) % 2 ** 31)
    x = q(rng.normal(size=(n, h, w, cin)).astype(np.float32), dt)
    wt = (rng.normal(size=(k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32)
    lib = N.lib()
    old = lib.dnnca_debug_force_generic(1 if variant == 'generic' else 0)
    P = (lambda lo, hi: (lo, hi)) if variant == 'sliced' else ((lambda lo, hi: (8 * lo, 8 * hi)) if variant == 'umma_sliced' else (lambda lo, hi: (0, 0)))
    if variant == 'umma_sliced':
        variant = 'umma'
        sliced = True
    else:
        sliced = variant == 'sliced'
    WS = (None, 0)
    if variant == 'umma':
        ws = torch.zeros(int(lib.dnnca_conv_workspace_bytes(k * k, cin, cout)), dtype=torch.uint8, device='cuda')
        _KEEP.append(ws)
        WS = (N.ptr(ws), ws.numel())
    try:
        wd, bd = dev(wt), dev(b)
        xa_b, xa_o, _ = embed(x[..., :ca], dt, *P(2, 1))
        xav = view(N, xa_b, xa_o, ca)
        xbv = None
        if cb:
            xb_b, xb_o, _ = embed(x[..., ca:], dt, *P(0, 3))
            xbv = view(N, xb_b, xb_o, cb)
        xbp = C.byref(xbv) if cb else None
        for act, alpha, tact in [(N.ACT_RELU, 0.0, 'relu'), (N.ACT_LEAKY, 0.3, ('leaky', 0.3)), (N.ACT_NONE, 0.0, None)]:
            lo, hi = P(1, 2)
            yb = torch.full((n, h, w, lo + cout + hi), 5.0, dtype=dt, device='cuda')
            stats = torch.zeros(2 * cout, dtype=torch.float64, device='cuda')
            yv = view(N, yb, lo, cout)
            N.call('dnnca_conv2d_fprop', None, C.byref(xav), xbp, N.ptr(wd), N.ptr(bd), C.byref(yv), k, act, alpha,
                   N.ptr(stats), *WS)
            sync()
            ref = ops.activation(ops.conv2d(torch.from_numpy(x), torch.from_numpy(wt), torch.from_numpy(b)), tact).numpy()
            got = yb[..., lo:lo + cout].float().cpu().numpy()
            close(got, ref, mode)
            if sliced:
                assert (yb[..., :lo].float() == 5.0).all() and (yb[..., lo + cout:].float() == 5.0).all(), 'wrote outside the slice'
            st = stats.cpu().numpy()
            np.testing.assert_allclose(st[:cout], got.astype(np.float64).sum((0, 1, 2)), rtol=1e-5, atol=1e-3)
            np.testing.assert_allclose(st[cout:], (got.astype(np.float64) ** 2).sum((0, 1, 2)), rtol=1e-5, atol=1e-3)
        # backward: dx is masked by relu'(mask); dx2 (skip gradient) is never masked
        dz = q(rng.normal(size=(n, h, w, cout)).astype(np.float32), dt)
        mask = q(rng.normal(size=(n, h, w, ca)).astype(np.float32), dt)
        dzb, dzo, _ = embed(dz, dt, *P(1, 2))
        mb, mo, _ = embed(mask, dt, *P(0, 1))
        lo, hi = P(2, 0)
        dxb = torch.full((n, h, w, lo + ca + hi), 3.0, dtype=dt, device='cuda')
        dx2b = torch.full((n, h, w, max(cb, 1) + lo), 3.0, dtype=dt, device='cuda')
        dzv, mv, dxv = view(N, dzb, dzo, cout), view(N, mb, mo, ca), view(N, dxb, lo, ca)
        dx2v = view(N, dx2b, lo, max(cb, 1))
        dx2p = C.byref(dx2v) if cb else None
        N.call('dnnca_conv2d_dgrad', None, C.byref(dzv), N.ptr(wd), C.byref(dxv), dx2p, k, C.byref(mv), N.ACT_RELU, 0.0, *WS)
        dw = torch.zeros(k, k, cin, cout, dtype=torch.float32, device='cuda')
        db = torch.zeros(cout, dtype=torch.float32, device='cuda')
        N.call('dnnca_conv2d_wgrad', None, C.byref(xav), xbp, C.byref(dzv), N.ptr(dw), N.ptr(db), k)
        sync()
        rdx, rdw, rdb = rn.conv2d_same_bwd(x, wt, dz)
        close(dxb[..., lo:lo + ca].float().cpu().numpy(), rdx[..., :ca] * (mask > 0), mode, scale=np.abs(rdx).max())
        if cb:
            close(dx2b[..., lo:].float().cpu().numpy(), rdx[..., ca:], mode, scale=np.abs(rdx).max())
        if sliced:
            assert (dxb[..., :lo].float() == 3.0).all()
        close(dw.cpu().numpy(), rdw, 'fp32', scale=np.abs(rdw).max() * (1 if mode == 'fp32' else 50))
        close(db.cpu().numpy(), rdb, 'fp32', scale=np.abs(rdb).max() * (1 if mode == 'fp32' else 50))
        # unmasked dgrad
        N.call('dnnca_conv2d_dgrad', None, C.byref(dzv), N.ptr(wd), C.byref(dxv), dx2p, k, None, N.ACT_NONE, 0.0, *WS)
        sync()
        close(dxb[..., lo:lo + ca].float().cpu().numpy(), rdx[..., :ca], mode, scale=np.abs(rdx).max())
    finally:
        lib.dnnca_debug_force_generic(old)


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('variant', ['dense', 'sliced'])
@pytest.mark.parametrize('shape', [(2, 8, 8, 12, 12), (1, 6, 16, 24, 8), (2, 5, 7, 6, 3), (1, 4, 4, 40, 70),
                                   (1, 20, 72, 12, 6), (2, 12, 40, 6, 3), (1, 36, 32, 8, 8),
                                   (1, 8, 16, 64, 32), (2, 16, 16, 128, 64), (1, 8, 32, 32, 16), (1, 8, 16, 384, 128),
                                   (1, 24, 20, 64, 64), (1, 16, 16, 512, 256), (2, 40, 24, 256, 64), (3, 12, 36, 128, 192)])
def test_tconv(N, mode, variant, shape):
    n, h, w, cin, cout = shape
    dt = DT[mode]
    rng = np.random.default_rng(sum(shape))
    if variant == 'dense':
        return _tconv_dense(N, mode, shape, rng)
    x = q(rng.normal(size=(n, h, w, cin)).astype(np.float32), dt)
    kt = (rng.normal(size=(2, 2, cout, cin)) / np.sqrt(cin)).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32)
    xb, xo, _ = embed(x, dt, 1, 1)
    yb = torch.full((n, 2 * h, 2 * w, 2 * cout), 5.0, dtype=dt, device='cuda')   # concat buffer, tconv half first
    stats = torch.zeros(2 * cout, dtype=torch.float64, device='cuda')
    xv, yv = view(N, xb, xo, cin), view(N, yb, 0, cout)
    N.call('dnnca_convtranspose2x2_fprop', None, C.byref(xv), N.ptr(dev(kt)), N.ptr(dev(b)), C.byref(yv), N.ptr(stats), None, 0)
    sync()
    ref = rn.tconv2x2_fwd(x, kt, b)
    got = yb[..., :cout].float().cpu().numpy()
    close(got, ref, mode)
    assert (yb[..., cout:].float() == 5.0).all()
    np.testing.assert_allclose(stats.cpu().numpy()[:cout], got.astype(np.float64).sum((0, 1, 2)), rtol=1e-5, atol=1e-3)
    dy = q(rng.normal(size=(n, 2 * h, 2 * w, cout)).astype(np.float32), dt)
    mask = q(rng.normal(size=(n, h, w, cin)).astype(np.float32), dt)
    dyb, dyo, _ = embed(dy, dt, 0, cout)
    mb, mo, _ = embed(mask, dt)
    dxb = torch.zeros(n, h, w, cin, dtype=dt, device='cuda')
    dyv, mv, dxv = view(N, dyb, dyo, cout), view(N, mb, mo, cin), view(N, dxb, 0, cin)
    N.call('dnnca_convtranspose2x2_dgrad', None, C.byref(dyv), N.ptr(dev(kt)), C.byref(dxv), C.byref(mv), N.ACT_LEAKY, 0.3, None, 0)
    dk = torch.zeros(2, 2, cout, cin, dtype=torch.float32, device='cuda')
    db = torch.zeros(cout, dtype=torch.float32, device='cuda')
    N.call('dnnca_convtranspose2x2_wgrad', None, C.byref(xv), C.byref(dyv), N.ptr(dk), N.ptr(db))
    sync()
    rdx, rdk, rdb = rn.tconv2x2_bwd(x, kt, dy)
    rdx = rn.act_bwd(mask, rdx, ('leaky', 0.3))
    close(dxb.float().cpu().numpy(), rdx, mode)
    close(dk.cpu().numpy(), rdk, 'fp32', scale=np.abs(rdk).max() * (1 if mode == 'fp32' else 50))
    close(db.cpu().numpy(), rdb, 'fp32', scale=np.abs(rdb).max() * (1 if mode == 'fp32' else 50))


UMMA_SHAPES = [  # (n, h, w, c_x, c_x2, cout, k): channel counts the tcgen05 implicit-GEMM kernels take
    (1, 16, 32, 64, 0, 64, 3), (2, 8, 16, 16, 0, 32, 3), (1, 16, 16, 64, 64, 64, 3), (1, 8, 16, 128, 0, 256, 3),
    (1, 24, 48, 32, 32, 32, 3), (2, 16, 16, 16, 16, 16, 3), (1, 8, 32, 256, 0, 128, 3), (1, 12, 20, 64, 0, 48, 3),
    (1, 16, 16, 64, 0, 64, 1), (1, 8, 16, 512, 512, 512, 3),
    # 32-channel layers (mulmo_unet.yaml) on the persistent halo kernel: half-empty N tile / zero-filled K chunk
    (2, 32, 32, 32, 0, 32, 3), (1, 16, 24, 16, 0, 32, 3), (1, 16, 16, 32, 0, 64, 3), (1, 16, 16, 64, 0, 32, 3), (1, 16, 16, 64, 0, 96, 3),
]


@pytest.mark.parametrize('variant', ['umma', 'umma_sliced'])
@pytest.mark.parametrize('shape', UMMA_SHAPES)
def test_conv_umma(N, variant, shape):
    """tcgen05/TMEM/TMA implicit-GEMM fprop + dgrad (bf16), dense and channel-sliced views; wgrad rides along."""
    lib = N.lib()
    lib.dnnca_debug_family_count(2, 1)
    lib.dnnca_debug_family_count(0, 1)
    _conv_case(N, 'bf16', variant, shape)
    wgrad_tc = True
    # 3 fprop + 2 dgrad (+ wgrad when the image tiles evenly) ran on the tensor cores, nothing fell back
    assert lib.dnnca_debug_family_count(2, 0) == 5 + (1 if wgrad_tc else 0)
    assert lib.dnnca_debug_family_count(0, 0) == (0 if wgrad_tc else 1)


def _tconv_dense(N, mode, shape, rng):
    """every tensor a whole dense buffer: the TMA-staged small-channel kernels take the shapes they cover"""
    n, h, w, cin, cout = shape
    dt = DT[mode]
    x = q(rng.normal(size=(n, h, w, cin)).astype(np.float32), dt)
    kt = (rng.normal(size=(2, 2, cout, cin)) / np.sqrt(cin)).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32)
    kd, bd = dev(kt), dev(b)
    WS = (None, 0)
    if mode == 'bf16' and cin % 16 == 0 and cout % 16 == 0:      # tensor-core path
        ws = torch.zeros(int(N.lib().dnnca_conv_workspace_bytes(4, cin, cout)), dtype=torch.uint8, device='cuda')
        _KEEP.append(ws)
        WS = (N.ptr(ws), ws.numel())
    xb, _, _ = embed(x, dt)
    yb = torch.full((n, 2 * h, 2 * w, cout), 5.0, dtype=dt, device='cuda')
    xv, yv = view(N, xb, 0, cin), view(N, yb, 0, cout)
    N.call('dnnca_convtranspose2x2_fprop', None, C.byref(xv), N.ptr(kd), N.ptr(bd), C.byref(yv), None, *WS)
    sync()
    close(yb.float().cpu().numpy(), rn.tconv2x2_fwd(x, kt, b), mode)
    dy = q(rng.normal(size=(n, 2 * h, 2 * w, cout)).astype(np.float32), dt)
    mask = q(rng.normal(size=(n, h, w, cin)).astype(np.float32), dt)
    dyb, _, _ = embed(dy, dt)
    mb, _, _ = embed(mask, dt)
    dxb = torch.full((n, h, w, cin), 9.0, dtype=dt, device='cuda')
    dyv, mv, dxv = view(N, dyb, 0, cout), view(N, mb, 0, cin), view(N, dxb, 0, cin)
    N.call('dnnca_convtranspose2x2_dgrad', None, C.byref(dyv), N.ptr(kd), C.byref(dxv), C.byref(mv), N.ACT_RELU, 0.0, *WS)
    dk = torch.zeros(2, 2, cout, cin, dtype=torch.float32, device='cuda')
    db = torch.zeros(cout, dtype=torch.float32, device='cuda')
    N.call('dnnca_convtranspose2x2_wgrad', None, C.byref(xv), C.byref(dyv), N.ptr(dk), N.ptr(db))
    sync()
    rdx, rdk, rdb = rn.tconv2x2_bwd(x, kt, dy)
    close(dxb.float().cpu().numpy(), rdx * (mask > 0), mode, scale=np.abs(rdx).max())
    close(dk.cpu().numpy(), rdk, 'fp32', scale=np.abs(rdk).max() * (1 if mode == 'fp32' else 50))
    close(db.cpu().numpy(), rdb, 'fp32', scale=np.abs(rdb).max() * (1 if mode == 'fp32' else 50))
    N.call('dnnca_convtranspose2x2_dgrad', None, C.byref(dyv), N.ptr(kd), C.byref(dxv), None, N.ACT_NONE, 0.0, *WS)
    sync()
    close(dxb.float().cpu().numpy(), rdx, mode)


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('c', [1, 3, 12, 64, 300])
def test_maxpool_fwd_bwd_bit_exact(N, mode, c):
    dt = DT[mode]
    rng = np.random.default_rng(c)
    n, h, w = 2, 12, 20
    x = q(np.maximum(rng.normal(size=(n, h, w, c)), 0).astype(np.float32), dt)   # relu output: many exact ties at 0
    xb, xo, _ = embed(x, dt, 1, 2)
    yb = torch.zeros(n, h // 2, w // 2, c + 1, dtype=dt, device='cuda')
    idx = torch.zeros(n, h // 2, w // 2, c, dtype=torch.uint8, device='cuda')
    stats = torch.zeros(2 * c, dtype=torch.float64, device='cuda')
    xv, yv = view(N, xb, xo, c), view(N, yb, 1, c)
    N.call('dnnca_maxpool2x2_fwd', None, C.byref(xv), C.byref(yv), N.ptr(idx), N.ptr(stats))
    sync()
    ry, ridx = rn.maxpool2x2_fwd(x)
    np.testing.assert_array_equal(yb[..., 1:].float().cpu().numpy(), ry)
    np.testing.assert_array_equal(idx.cpu().numpy(), ridx)          # first-max-wins, bit exact
    ty, tidx = ops.maxpool(torch.from_numpy(x), 2, return_indices=True)
    np.testing.assert_array_equal(idx.cpu().numpy(), tidx.numpy())
    np.testing.assert_allclose(stats.cpu().numpy()[:c], ry.astype(np.float64).sum((0, 1, 2)), rtol=1e-6, atol=1e-4)
    # backward with skip gradient added in place and relu mask from x
    dy = q(rng.normal(size=ry.shape).astype(np.float32), dt)
    dskip = q(rng.normal(size=x.shape).astype(np.float32), dt)
    dyb, dyo, _ = embed(dy, dt)
    dxb, dxo, _ = embed(dskip, dt, 3, 0)
    dyv, dxv = view(N, dyb, dyo, c), view(N, dxb, dxo, c)
    N.call('dnnca_maxpool2x2_bwd', None, C.byref(dyv), N.ptr(idx), C.byref(dxv), C.byref(dxv), C.byref(xv), N.ACT_RELU, 0.0)
    sync()
    ref = (rn.maxpool2x2_bwd(dy, ridx) + dskip) * (x > 0)
    close(dxb[..., 3:].float().cpu().numpy(), ref, mode)
    # plain scatter
    N.call('dnnca_maxpool2x2_bwd', None, C.byref(dyv), N.ptr(idx), None, C.byref(dxv), None, N.ACT_NONE, 0.0)
    sync()
    np.testing.assert_array_equal(dxb[..., 3:].float().cpu().numpy(), rn.maxpool2x2_bwd(dy, ridx))


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('c,scale', [(3, True), (16, True), (64, False), (130, True), (600, True)])
def test_batchnorm_train_fwd_bwd_and_inference(N, mode, c, scale):
    dt = DT[mode]
    rng = np.random.default_rng(c)
    n, h, w = 2, 10, 12
    x = q(np.maximum(rng.normal(0.3, 1.0, size=(n, h, w, c)), 0).astype(np.float32), dt)
    gamma = rng.uniform(.5, 1.5, c).astype(np.float32) if scale else None
    beta = rng.normal(0, .1, c).astype(np.float32)
    mm0, mv0 = rng.normal(0, .1, c).astype(np.float32), rng.uniform(.5, 1.5, c).astype(np.float32)
    pad = 8 if c % 8 == 0 else 1          # multiples of 8 keep 16-byte alignment: the 128-bit vectorised bf16 kernels
    xb, xo, _ = embed(x, dt, pad, pad)
    xv = view(N, xb, xo, c)
    stats = torch.zeros(4 * c, dtype=torch.float64, device='cuda')
    N.call('dnnca_channel_stats', None, C.byref(xv), N.ptr(stats))
    ss = torch.zeros(2 * c, device='cuda')
    mi = torch.zeros(2 * c, device='cuda')
    mm, mv = dev(mm0), dev(mv0)
    g = dev(gamma) if scale else None
    N.call('dnnca_bn_finalize', None, N.ptr(stats), n * h * w, c, N.ptr(g), N.ptr(dev(beta)), 0.99, 1e-3, N.ptr(mm),
           N.ptr(mv), N.ptr(ss), N.ptr(mi))
    yo = 2 * pad
    yb = torch.zeros(n, h, w, c + yo, dtype=dt, device='cuda')
    yv = view(N, yb, yo, c)
    N.call('dnnca_bn_apply', None, C.byref(xv), N.ptr(ss), C.byref(yv))
    sync()
    T = lambda a: torch.tensor(a, dtype=torch.float64)
    ry, rmm, rmv = ops.batchnorm(T(x), T(gamma) if scale else None, T(beta), T(mm0), T(mv0), True)
    close(yb[..., yo:].float().cpu().numpy(), ry.numpy(), mode)
    np.testing.assert_allclose(mm.cpu().numpy(), rmm.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mv.cpu().numpy(), rmv.numpy(), rtol=1e-5, atol=1e-6)
    # backward (x is a relu output feeding the BN -> fused relu mask)
    dy = q(rng.normal(size=x.shape).astype(np.float32), dt)
    dyb, dyo, _ = embed(dy, dt, 0, 3 * pad)
    dyv = view(N, dyb, dyo, c)
    dxb = torch.zeros(n, h, w, c, dtype=dt, device='cuda')
    dxv = view(N, dxb, 0, c)
    sums = stats[2 * c:]
    N.call('dnnca_bn_bwd_reduce', None, C.byref(xv), C.byref(dyv), N.ptr(mi), N.ptr(sums))
    dg, dbt = torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
    N.call('dnnca_bn_bwd_apply', None, C.byref(xv), C.byref(dyv), N.ptr(mi), N.ptr(g), N.ptr(sums), C.byref(dxv),
           N.ACT_RELU, 0.0, N.ptr(dg) if scale else None, N.ptr(dbt))
    sync()
    _, mean, var, invstd = rn.bn_train_fwd(x, gamma, beta)
    rdx, rdg, rdb = rn.bn_train_bwd(x, gamma, dy, mean, invstd)
    close(dxb.float().cpu().numpy(), rdx * (x > 0), mode)
    np.testing.assert_allclose(dbt.cpu().numpy(), rdb, rtol=1e-4, atol=1e-4)
    if scale:
        np.testing.assert_allclose(dg.cpu().numpy(), rdg, rtol=1e-4, atol=1e-3)
    # inference parameters from the moving statistics
    N.call('dnnca_bn_inference_params', None, c, N.ptr(g), N.ptr(dev(beta)), 1e-3, N.ptr(dev(mm0)), N.ptr(dev(mv0)), N.ptr(ss))
    N.call('dnnca_bn_apply', None, C.byref(xv), N.ptr(ss), C.byref(yv))
    sync()
    ry, _, _ = ops.batchnorm(T(x), T(gamma) if scale else None, T(beta), T(mm0), T(mv0), False)
    close(yb[..., yo:].float().cpu().numpy(), ry.numpy(), mode)


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('F,healthy,weight', [(3, False, None), (16, False, None), (64, True, None), (3, False, 5.0)])
def test_head_bce_fused(N, mode, F, healthy, weight):
    _head_case(N, mode, F, healthy, weight, 8 if F % 8 == 0 else 1)


@pytest.mark.parametrize('F,healthy,weight', [(3, False, None), (3, True, None), (4, False, 2.0), (6, False, None)])
def test_head_bce_pix8_dense(N, F, healthy, weight):
    """dense bf16 features with few channels (configs/unet.yaml: F = 3) take the 8-pixels-per-thread kernel"""
    _head_case(N, 'bf16', F, healthy, weight, 0)


def _head_case(N, mode, F, healthy, weight, pad):
    dt = DT[mode]
    rng = np.random.default_rng(F)
    n, h, w = 3, 16, 24
    f = q(np.maximum(rng.normal(size=(n, h, w, F)), 0).astype(np.float32), dt)
    wt = rng.normal(size=F).astype(np.float32)
    b = np.array([0.1], np.float32)
    y = (rng.uniform(size=(n, h, w)) < (0.0 if healthy else 0.05)).astype(np.float32)
    fb, fo, _ = embed(f, dt, pad, 0)      # 16-byte aligned slices take the vectorised bf16 kernel
    fv = view(N, fb, fo, F)
    ls = torch.zeros(16, dtype=torch.uint8, device='cuda')
    yd = dev(y)
    N.call('dnnca_label_stats_init', None, N.ptr(ls))
    N.call('dnnca_label_stats', None, N.ptr(yd), y.size, N.ptr(ls))
    cfg = N.LossConfig(weight or 0.0, 1 if weight is not None else 0, 0.0, 3.0, 1.0 / (y.size * 2))
    logits, probs = torch.zeros(n, h, w, device='cuda'), torch.zeros(n, h, w, device='cuda')
    per = torch.zeros(n, device='cuda')
    dfb = torch.zeros(n, h, w, F, dtype=dt, device='cuda')
    dfv = view(N, dfb, 0, F)
    dw, db = torch.zeros(F, device='cuda'), torch.zeros(1, device='cuda')
    N.call('dnnca_head_bce_fwd_bwd', None, C.byref(fv), N.ptr(dev(wt)), N.ptr(dev(b)), N.ptr(yd), N.ptr(ls),
           C.byref(cfg), N.ptr(logits), N.ptr(probs), N.ptr(per), C.byref(dfv), N.ACT_RELU, 0.0, N.ptr(dw), N.ptr(db))
    sync()
    host = N.LabelStats.from_buffer_copy(ls.cpu().numpy().tobytes())
    s, mn, mx = C.c_double(), C.c_float(), C.c_float()
    N.lib().dnnca_label_stats_decode(C.byref(host), C.byref(s), C.byref(mn), C.byref(mx))
    assert abs(s.value - y.sum()) < 1e-6 and mn.value == y.min() and mx.value == y.max()
    z = f.astype(np.float64) @ wt.astype(np.float64) + b[0]
    rper, rdz = rn.weighted_bce_fwd_bwd(y, z, weight=weight, weight_mul=3.0, n_replicas=2)
    np.testing.assert_allclose(logits.cpu().numpy(), z, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(probs.cpu().numpy(), 1 / (1 + np.exp(-z)), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(per.cpu().numpy(), rper, rtol=1e-4)
    rdf = rdz[..., None] * wt * (f > 0)
    close(dfb.float().cpu().numpy(), rdf, mode)
    np.testing.assert_allclose(dw.cpu().numpy(), np.einsum('nhw,nhwf->f', rdz, f.astype(np.float64)), rtol=1e-3, atol=1e-7)
    np.testing.assert_allclose(db.cpu().numpy()[0], rdz.sum(), rtol=1e-3, atol=1e-8)
    # forward-only head
    l2, p2 = torch.zeros_like(logits), torch.zeros_like(probs)
    N.call('dnnca_head_fwd', None, C.byref(fv), N.ptr(dev(wt)), N.ptr(dev(b)), N.ptr(l2), N.ptr(p2))
    sync()
    np.testing.assert_allclose(l2.cpu().numpy(), z, rtol=1e-5, atol=1e-5)


def test_adam_matches_keras_form(N):
    rng = np.random.default_rng(0)
    n = 1000
    p0, g = rng.normal(size=n).astype(np.float32), rng.normal(size=n).astype(np.float32)
    p, m, v = dev(p0), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    hyper = dev(np.array([1e-3, 0.9, 0.999, 1e-7], np.float32))
    step = torch.zeros(1, dtype=torch.int64, device='cuda')
    rp, rm, rv = torch.tensor(p0, dtype=torch.float64), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    for t in range(1, 4):
        N.call('dnnca_adam_step', None, N.ptr(p), N.ptr(dev(g)), N.ptr(m), N.ptr(v), n, N.ptr(hyper), N.ptr(step), None)
        rp, rm, rv = ops.adam_step(rp, torch.tensor(g, dtype=torch.float64), rm, rv, t)
    sync()
    assert int(step) == 3
    np.testing.assert_allclose(p.cpu().numpy(), rp.numpy(), rtol=1e-5, atol=1e-6)
    # L2 term: g + 2*l2*p
    l2 = dev(np.full(n, 0.01, np.float32))
    p2, m2, v2 = dev(p0), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    step.zero_()
    N.call('dnnca_adam_step', None, N.ptr(p2), N.ptr(dev(g)), N.ptr(m2), N.ptr(v2), n, N.ptr(hyper), N.ptr(step), N.ptr(l2))
    sync()
    rp2, _, _ = ops.adam_step(torch.tensor(p0, dtype=torch.float64), torch.tensor(g + 0.02 * p0, dtype=torch.float64),
                              torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64), 1)
    np.testing.assert_allclose(p2.cpu().numpy(), rp2.numpy(), rtol=1e-5, atol=1e-6)


def test_u8_to_unit_and_convert_bit_exact(N):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, 10007, dtype=np.uint8)
    src = torch.from_numpy(a).cuda()
    dst = torch.zeros(a.size, device='cuda')
    N.call('dnnca_u8_to_unit', None, N.ptr(src), a.size, N.ptr(dst), N.F32)
    sync()
    np.testing.assert_array_equal(dst.cpu().numpy(), a.astype(np.float32) / np.float32(255.0))   # data.py:206
    x = rng.normal(size=(2, 5, 7, 6)).astype(np.float32)
    xb, xo, _ = embed(x, torch.float32, 1, 1)
    yb = torch.zeros(2, 5, 7, 9, dtype=torch.bfloat16, device='cuda')
    xv, yv = view(N, xb, xo, 6), view(N, yb, 3, 6)
    N.call('dnnca_convert', None, C.byref(xv), C.byref(yv))
    sync()
    np.testing.assert_array_equal(yb[..., 3:].float().cpu().numpy(), q(x, torch.bfloat16))


def test_add_relu_affine(N):
    rng = np.random.default_rng(1)
    c = 51
    a, b = rng.normal(size=(2, 6, 6, c)).astype(np.float32), rng.normal(size=(2, 6, 6, c)).astype(np.float32)
    fb, fo = rng.normal(size=2 * c).astype(np.float32), rng.normal(size=2 * c).astype(np.float32)
    ab, ao, _ = embed(a, torch.float32, 1, 0)
    bb, bo, _ = embed(b, torch.float32)
    yb = torch.zeros(2, 6, 6, c, device='cuda')
    av, bv, yv = view(N, ab, ao, c), view(N, bb, bo, c), view(N, yb, 0, c)
    N.call('dnnca_add_relu_affine', None, C.byref(av), None, C.byref(bv), N.ptr(dev(fb)), N.ptr(dev(fo)), C.byref(yv))
    sync()
    ref = np.maximum(a + b * fb[:c] + fb[c:], 0) * fo[:c] + fo[c:]
    np.testing.assert_allclose(yb.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('c,pads', [(32, (0, 0)), (120, (8, 16)), (216, (0, 8)), (864, (0, 0)), (2056, (0, 0))])
def test_add_relu_affine_bf16_vector_paths(N, c, pads):
    """bf16 views with 8-channel alignment: the register-table kernel (c / 8 <= 256 groups, incl. the MultiRes widths that
    are no power of two) and the shared-memory one beyond; ragged pixel counts; every affine present / absent."""
    rng = np.random.default_rng(c)
    shape = (2, 5, 7, c)
    a, b = q(rng.normal(size=shape).astype(np.float32), torch.bfloat16), q(rng.normal(size=shape).astype(np.float32), torch.bfloat16)
    fa, fb, fo = (rng.normal(size=2 * c).astype(np.float32) for _ in range(3))
    ab, ao, _ = embed(a, torch.bfloat16, *pads)
    bb, bo, _ = embed(b, torch.bfloat16)
    for use in ((True, True, True), (False, True, False), (False, False, True)):
        yb = torch.zeros(*shape, dtype=torch.bfloat16, device='cuda')
        av, bv, yv = view(N, ab, ao, c), view(N, bb, bo, c), view(N, yb, 0, c)
        N.call('dnnca_add_relu_affine', None, C.byref(av), N.ptr(dev(fa)) if use[0] else None, C.byref(bv),
               N.ptr(dev(fb)) if use[1] else None, N.ptr(dev(fo)) if use[2] else None, C.byref(yv))
        sync()
        va = a * fa[:c] + fa[c:] if use[0] else a
        vb = b * fb[:c] + fb[c:] if use[1] else b
        ref = np.maximum(va + vb, 0)
        ref = ref * fo[:c] + fo[c:] if use[2] else ref
        close(yb.float().cpu().numpy(), ref, 'bf16')
        assert float(ab[..., :ao].float().abs().min()) == 7.0 if ao else True          # neighbours of the view untouched


@pytest.mark.parametrize('F', [8, 16, 64, 256, 24])
def test_head_fwd_bf16_vector_path(N, F):
    """dnnca_head_fwd on bf16 features: F = 8 * 2^k takes the 16-byte / shuffle kernel (F = 24 stays on the scalar one);
    35 pixels = a ragged last warp."""
    rng = np.random.default_rng(F)
    f = q(rng.normal(size=(1, 5, 7, F)).astype(np.float32), torch.bfloat16)
    wt, b = rng.normal(size=F).astype(np.float32), rng.normal(size=1).astype(np.float32)
    fbuf, fo_, _ = embed(f, torch.bfloat16, 8, 8)
    fv = view(N, fbuf, fo_, F)
    lg, pr = torch.zeros(1, 5, 7, device='cuda'), torch.zeros(1, 5, 7, device='cuda')
    N.call('dnnca_head_fwd', None, C.byref(fv), N.ptr(dev(wt)), N.ptr(dev(b)), N.ptr(lg), N.ptr(pr))
    sync()
    z = f.astype(np.float64) @ wt.astype(np.float64) + b[0]
    np.testing.assert_allclose(lg.cpu().numpy(), z, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(pr.cpu().numpy(), 1 / (1 + np.exp(-z)), rtol=1e-5, atol=1e-6)
    only = torch.zeros(1, 5, 7, device='cuda')
    N.call('dnnca_head_fwd', None, C.byref(fv), N.ptr(dev(wt)), None, N.ptr(only), None)      # no bias, logits only
    sync()
    np.testing.assert_allclose(only.cpu().numpy(), z - b[0], rtol=1e-5, atol=1e-5)


def test_bad_arguments_fail_loudly(N):
    x = torch.zeros(1, 4, 4, 3, device='cuda')
    xv = N.tensor_view(x)
    yv = N.tensor_view(torch.zeros(1, 4, 4, 3, device='cuda'))
    with pytest.raises(N.DnncaError, match='kernel size'):
        N.call('dnnca_conv2d_fprop', None, C.byref(xv), None, N.ptr(x), None, C.byref(yv), 5, 0, 0.0, None, None, 0)
    yv2 = N.tensor_view(torch.zeros(1, 5, 4, 3, device='cuda'))
    with pytest.raises(N.DnncaError):
        N.call('dnnca_conv2d_fprop', None, C.byref(xv), None, N.ptr(x), None, C.byref(yv2), 3, 0, 0.0, None, None, 0)


@pytest.mark.parametrize('shape', [(4, 128, 128, 64, 0, 64), (2, 96, 72, 64, 64, 128), (3, 64, 64, 128, 0, 256),
                                   (2, 40, 56, 256, 0, 64)])
def test_conv_umma_persistent_many_tiles(N, shape):
    """More pixel tiles than persistent CTAs: exercises the halo-tile ring, the resident / streamed weight modes and
    the double-buffered TMEM hand-off of the second-generation tcgen05 kernel over several tiles per CTA
    (fprop with two inputs, dgrad with two outputs and mask), against torch autograd."""
    n, h, w, ca, cb, cout = shape
    cin = ca + cb
    dt = torch.bfloat16
    rng = np.random.default_rng(sum(shape))
    x = q(rng.normal(size=(n, h, w, cin)).astype(np.float32), dt)
    wt = (rng.normal(size=(3, 3, cin, cout)) / np.sqrt(9 * cin)).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32)
    dz = q(rng.normal(size=(n, h, w, cout)).astype(np.float32), dt)
    lib = N.lib()
    ws = torch.zeros(int(lib.dnnca_conv_workspace_bytes(9, cin, cout)), dtype=torch.uint8, device='cuda')
    wd, bd = dev(wt), dev(b)
    xa = dev(x[..., :ca], dt)
    xb = dev(x[..., ca:], dt) if cb else None
    y = torch.zeros(n, h, w, cout, dtype=dt, device='cuda')
    xav, yv = N.tensor_view(xa), N.tensor_view(y)
    xbv = N.tensor_view(xb) if cb else None
    lib.dnnca_debug_family_count(2, 1)
    N.call('dnnca_conv2d_fprop', None, C.byref(xav), C.byref(xbv) if cb else None, N.ptr(wd), N.ptr(bd), C.byref(yv), 3,
           N.ACT_RELU, 0.0, None, N.ptr(ws), ws.numel())
    sync()
    xt = torch.from_numpy(x).requires_grad_()
    ref = torch.relu(ops.conv2d(xt, torch.from_numpy(wt), torch.from_numpy(b)))
    close(y.float().cpu().numpy(), ref.detach().numpy(), 'bf16')
    # dgrad: two outputs, relu mask on the first
    dzd = dev(dz, dt)
    dxa = torch.zeros(n, h, w, ca, dtype=dt, device='cuda')
    dxb = torch.zeros(n, h, w, max(cb, 1), dtype=dt, device='cuda')
    dzv, dxav, dxbv = N.tensor_view(dzd), N.tensor_view(dxa), N.tensor_view(dxb)
    N.call('dnnca_conv2d_dgrad', None, C.byref(dzv), N.ptr(wd), C.byref(dxav), C.byref(dxbv) if cb else None, 3,
           C.byref(xav), N.ACT_RELU, 0.0, N.ptr(ws), ws.numel())
    sync()
    lin = ops.conv2d(xt, torch.from_numpy(wt), None)
    (g,) = torch.autograd.grad(lin, xt, torch.from_numpy(dz))
    g = g.numpy()
    close(dxa.float().cpu().numpy(), g[..., :ca] * (x[..., :ca] > 0), 'bf16', scale=np.abs(g).max())
    if cb:
        close(dxb.float().cpu().numpy(), g[..., ca:], 'bf16', scale=np.abs(g).max())
    assert lib.dnnca_debug_family_count(2, 0) == 2


# ---- row-Toeplitz tcgen05 kernels (conv_row_umma.cu): every conv shape of configs/unet.yaml, several tiles per CTA ------
ROW_SHAPES = [  # (n, h, w, c_x, c_x2, cout)
    (2, 256, 256, 3, 0, 3), (2, 128, 128, 3, 0, 6), (3, 128, 128, 6, 0, 6), (3, 64, 64, 6, 0, 12), (5, 64, 64, 12, 0, 12),
    (3, 64, 64, 12, 12, 12), (2, 128, 128, 6, 6, 6), (2, 256, 256, 3, 3, 3), (40, 32, 64, 3, 0, 3), (2, 48, 32, 6, 0, 6),
    (1, 16, 16, 4, 4, 4), (2, 32, 32, 5, 0, 3), (150, 128, 32, 3, 0, 3),
    # images of <= 64 rows are interleaved row by row, 2 / 4 / 8 per 128-row tile (even batch)
    (4, 64, 64, 12, 12, 12), (6, 64, 64, 6, 0, 12), (8, 16, 32, 6, 0, 6), (12, 32, 64, 12, 0, 6), (2, 64, 128, 3, 3, 3),
]


@pytest.mark.parametrize('shape', ROW_SHAPES)
def test_conv_row_umma(N, shape):
    """bf16 few-channel 3x3 convs on the tensor cores: fprop (+bias+ReLU), masked and unmasked dgrad with two
    destinations, wgrad + bias gradient, against torch fp64 autograd on the same bf16-rounded inputs."""
    n, h, w, ca, cb, cout = shape
    cin = ca + cb
    lib = N.lib()
    rng = np.random.default_rng(abs(hash(shape)) % 2 ** 31)
    bf = torch.bfloat16
    x = q(rng.normal(size=(n, h, w, cin)).astype(np.float32), bf)
    wt = (rng.normal(size=(3, 3, cin, cout)) / np.sqrt(9 * cin)).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32)
    dz = q(rng.normal(size=(n, h, w, cout)).astype(np.float32), bf)
    mask = q(rng.normal(size=(n, h, w, ca)).astype(np.float32), bf)
    wd, bd = dev(wt), dev(b)
    xa = dev(x[..., :ca], bf)
    xav = view(N, xa, 0, ca)
    xbp = None
    if cb:
        xb = dev(x[..., ca:], bf)
        xbv = view(N, xb, 0, cb)
        xbp = C.byref(xbv)
    y = torch.full((n, h, w, cout), 5.0, dtype=bf, device='cuda')
    yv = view(N, y, 0, cout)
    for fam in (0, 1, 2):
        lib.dnnca_debug_family_count(fam, 1)
    N.call('dnnca_conv2d_fprop', None, C.byref(xav), xbp, N.ptr(wd), N.ptr(bd), C.byref(yv), 3, N.ACT_RELU, 0.0, None, None, 0)
    dzd, md = dev(dz, bf), dev(mask, bf)
    dzv, mv = view(N, dzd, 0, cout), view(N, md, 0, ca)
    dx = torch.full((n, h, w, ca), 3.0, dtype=bf, device='cuda')
    dx2 = torch.full((n, h, w, max(cb, 1)), 3.0, dtype=bf, device='cuda')
    dxv, dx2v = view(N, dx, 0, ca), view(N, dx2, 0, max(cb, 1))
    dx2p = C.byref(dx2v) if cb else None
    N.call('dnnca_conv2d_dgrad', None, C.byref(dzv), N.ptr(wd), C.byref(dxv), dx2p, 3, C.byref(mv), N.ACT_RELU, 0.0, None, 0)
    dw = torch.zeros(3, 3, cin, cout, dtype=torch.float32, device='cuda')
    db = torch.zeros(cout, dtype=torch.float32, device='cuda')
    N.call('dnnca_conv2d_wgrad', None, C.byref(xav), xbp, C.byref(dzv), N.ptr(dw), N.ptr(db), 3)
    sync()
    assert lib.dnnca_debug_family_count(2, 0) == 3, 'the row-Toeplitz tcgen05 kernels did not take this shape'
    # reference: torch fp64 autograd (NCHW) on the same rounded inputs
    xt = torch.from_numpy(x).double().permute(0, 3, 1, 2).requires_grad_(True)
    wtt = torch.from_numpy(wt).double().permute(3, 2, 0, 1).requires_grad_(True)
    bt = torch.from_numpy(b).double().requires_grad_(True)
    pre = torch.nn.functional.conv2d(xt, wtt, bt, padding=1)
    pre.backward(torch.from_numpy(dz).double().permute(0, 3, 1, 2))
    ref_y = torch.relu(pre).detach().permute(0, 2, 3, 1).numpy()
    rdx = xt.grad.permute(0, 2, 3, 1).numpy()
    rdw = wtt.grad.permute(2, 3, 1, 0).numpy()
    rdb = bt.grad.numpy()
    close(y.float().cpu().numpy(), ref_y, 'bf16')
    close(dx.float().cpu().numpy(), rdx[..., :ca] * (mask > 0), 'bf16', scale=np.abs(rdx).max())
    if cb:
        close(dx2.float().cpu().numpy(), rdx[..., ca:], 'bf16', scale=np.abs(rdx).max())
    close(dw.cpu().numpy(), rdw, 'fp32', scale=np.abs(rdw).max() * 50)
    close(db.cpu().numpy(), rdb, 'fp32', scale=np.abs(rdb).max() * 50)
    N.call('dnnca_conv2d_dgrad', None, C.byref(dzv), N.ptr(wd), C.byref(dxv), dx2p, 3, None, N.ACT_NONE, 0.0, None, 0)
    sync()
    close(dx.float().cpu().numpy(), rdx[..., :ca], 'bf16', scale=np.abs(rdx).max())


@pytest.mark.parametrize('c,h,w', [(3, 12, 32), (6, 8, 16), (12, 20, 8), (3, 256, 256), (64, 8, 12), (16, 12, 20), (136, 6, 10)])
def test_maxpool_vec_dense_bit_exact(N, c, h, w):
    """dense bf16 tensors with 3/6/12 channels take the 128-bit vectorised kernels (pool_vec.cu); results must be
    bit-identical to the oracle (first maximum wins, skip gradient added, ReLU mask from the pooled layer's input)."""
    bf = torch.bfloat16
    rng = np.random.default_rng(c * 1000 + h)
    n = 2
    x = q(np.maximum(rng.normal(size=(n, h, w, c)), 0).astype(np.float32), bf)
    xd = dev(x, bf)
    y = torch.zeros(n, h // 2, w // 2, c, dtype=bf, device='cuda')
    idx = torch.zeros(n, h // 2, w // 2, c, dtype=torch.uint8, device='cuda')
    xv, yv = view(N, xd, 0, c), view(N, y, 0, c)
    stats = torch.zeros(2 * c, dtype=torch.float64, device='cuda') if c % 8 == 0 else None   # fused BN statistics (vec8 kernel)
    N.call('dnnca_maxpool2x2_fwd', None, C.byref(xv), C.byref(yv), N.ptr(idx), N.ptr(stats))
    sync()
    ty, tidx = ops.maxpool(torch.from_numpy(x), 2, return_indices=True)
    np.testing.assert_array_equal(y.float().cpu().numpy(), ty.numpy())
    np.testing.assert_array_equal(idx.cpu().numpy(), tidx.numpy())
    if stats is not None:
        t64 = ty.numpy().astype(np.float64)
        np.testing.assert_allclose(stats.cpu().numpy()[:c], t64.sum((0, 1, 2)), rtol=1e-6, atol=1e-4)
        np.testing.assert_allclose(stats.cpu().numpy()[c:], (t64 ** 2).sum((0, 1, 2)), rtol=1e-6, atol=1e-4)
    dy = q(rng.normal(size=ty.shape).astype(np.float32), bf)
    dskip = q(rng.normal(size=x.shape).astype(np.float32), bf)
    dyd, dxd = dev(dy, bf), dev(dskip, bf)
    dyv, dxv = view(N, dyd, 0, c), view(N, dxd, 0, c)
    N.call('dnnca_maxpool2x2_bwd', None, C.byref(dyv), N.ptr(idx), C.byref(dxv), C.byref(dxv), C.byref(xv), N.ACT_RELU, 0.0)
    sync()
    scat = np.zeros_like(x)
    ti = tidx.numpy()
    for a in range(2):
        for b in range(2):
            scat[:, a::2, b::2, :] = np.where(ti == 2 * a + b, dy, 0.0)
    ref = q(((scat + dskip) * (x > 0)).astype(np.float32), bf)
    np.testing.assert_array_equal(dxd.float().cpu().numpy(), ref)
    N.call('dnnca_maxpool2x2_bwd', None, C.byref(dyv), N.ptr(idx), None, C.byref(dxv), None, N.ACT_NONE, 0.0)
    sync()
    np.testing.assert_array_equal(dxd.float().cpu().numpy(), scat)
    # flat fp32 -> bf16 conversion (input staging)
    src = dev(rng.normal(size=(n, h, w, c)).astype(np.float32))
    dst = torch.zeros(n, h, w, c, dtype=bf, device='cuda')
    sv, dv = view(N, src, 0, c), view(N, dst, 0, c)
    N.call('dnnca_convert', None, C.byref(sv), C.byref(dv))
    sync()
    assert torch.equal(dst, src.to(bf))


@pytest.mark.parametrize('shape', [(3, 32, 32, 12, 12), (2, 64, 64, 12, 6), (2, 128, 128, 6, 3), (150, 16, 32, 6, 3),
                                   (2, 48, 16, 12, 6), (1, 64, 64, 3, 3)])
def test_tconv_row_umma(N, shape):
    """Conv2DTranspose k=s=2 with few channels (configs/unet.yaml decoder) on the row-Toeplitz tcgen05 kernels:
    fprop + bias, masked and unmasked dgrad, wgrad + bias gradient."""
    lib = N.lib()
    for fam in (0, 1, 2):
        lib.dnnca_debug_family_count(fam, 1)
    _tconv_dense(N, 'bf16', shape, np.random.default_rng(sum(shape)))
    assert lib.dnnca_debug_family_count(2, 0) == 4, 'the row-Toeplitz tcgen05 kernels did not take this ConvT shape'


@pytest.mark.parametrize('shape,size,sigma', [((3, 32, 40), 6, 3.0), ((2, 17, 9), 5, 1.5), ((1, 256, 256), 6, 3.0), ((4, 8, 8), 3, 0.7)])
def test_gaussian_label_smoothing(N, shape, size, sigma):
    """losses.py:62-67 (tfa.image.gaussian_filter2d, REFLECT padding) against the oracle restatement."""
    from oracle import ref_ops as ops
    rng = np.random.default_rng(sum(shape) + size)
    y = (rng.random(shape) > 0.7).astype(np.float32)
    yd = dev(y)
    tmp, out = torch.empty_like(yd), torch.empty_like(yd)
    N.call('dnnca_gaussian_filter2d', None, N.ptr(yd), shape[0], shape[1], shape[2], size, sigma, N.ptr(tmp), N.ptr(out))
    sync()
    ref = ops.gaussian_filter2d(torch.tensor(y), size, sigma).numpy()
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)
    assert abs(float(out.sum()) - float(ref.sum())) <= 1e-3 * max(1.0, float(ref.sum()))
