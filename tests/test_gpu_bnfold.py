"""BatchNormalization folded into its consumers (csrc/bn_fold.cu; components.py:46-61, 118-134 block order
Conv -> act -> BN -> Conv / MaxPool): per-op parity of the folded entry points against the oracle evaluated on the
materialised BN output, and model-level agreement of the folded plan with the unfolded one and with the fp32 oracle."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_models as rm
from oracle import ref_ops as ops

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def N():
    from dnncancerannotator_b200 import native
    native.lib()
    return native


def bf(a):
    return torch.as_tensor(a).bfloat16().float()


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# (n, h, w, cx, cx2, cout): single inputs of any width, two-input (virtual concat) layers, partial N tiles
FOLD_SHAPES = [(2, 32, 24, 64, 0, 64), (1, 16, 40, 32, 0, 64), (2, 18, 22, 64, 64, 64), (1, 34, 16, 128, 0, 128),
               (2, 16, 16, 40, 0, 24), (1, 32, 32, 128, 128, 128), (1, 20, 12, 64, 0, 256)]


@pytest.mark.parametrize('shape', FOLD_SHAPES)
@pytest.mark.parametrize('act', ['relu', 'leaky'])
def test_conv_fprop_and_wgrad_with_folded_input_affine(N, shape, act):
    n, h, w, cx, cx2, cout = shape
    rng = np.random.default_rng(abs(hash(shape)) % 2 ** 31)
    lib = N.lib()
    a1 = bf(np.maximum(rng.normal(0.3, 1.0, (n, h, w, cx)), 0).astype(np.float32))          # post-ReLU, like the real producer
    a2 = bf(np.maximum(rng.normal(0.1, 1.0, (n, h, w, cx2)), 0).astype(np.float32)) if cx2 else None
    aff1 = np.concatenate([rng.uniform(-1.5, 1.5, cx), rng.normal(0, 0.5, cx)]).astype(np.float32)      # negative scales too
    aff2 = np.concatenate([rng.uniform(0.5, 1.5, cx2), rng.normal(0, 0.5, cx2)]).astype(np.float32) if cx2 else None
    cin = cx + cx2
    wt = (rng.normal(0, 1.0, (3, 3, cin, cout)) / np.sqrt(9 * cin)).astype(np.float32)
    bias = rng.normal(0, 0.1, cout).astype(np.float32)
    dz = bf(rng.normal(0, 1.0, (n, h, w, cout)).astype(np.float32))
    # oracle on the MATERIALISED BN output (fp32, never rounded), zero padding after the affine
    x = a1 * torch.tensor(aff1[:cx]) + torch.tensor(aff1[cx:])
    if cx2:
        x = torch.cat([x, a2 * torch.tensor(aff2[:cx2]) + torch.tensor(aff2[cx2:])], -1)
    wq = torch.tensor(wt, requires_grad=True)
    actspec = 'relu' if act == 'relu' else ('leaky', 0.3)
    y_ref = ops.activation(ops.conv2d(x, wq, torch.tensor(bias), 'same'), actspec)
    pre = ops.conv2d(x, wq, torch.tensor(bias), 'same')
    gw, = torch.autograd.grad(pre, wq, dz)
    # device
    d = lambda t, dt=torch.bfloat16: None if t is None else torch.as_tensor(t).to('cuda').to(dt).contiguous()
    xa, xb, dzd = d(a1), d(a2), d(dz)
    fa, fb = d(aff1, torch.float32), d(aff2, torch.float32)
    wd, bd = d(wt, torch.float32), d(bias, torch.float32)
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device='cuda')
    va, vb, vy, vdz = N.tensor_view(xa), (N.tensor_view(xb) if cx2 else None), N.tensor_view(y), N.tensor_view(dzd)
    assert lib.dnnca_conv2d_fold_supported(C.byref(va), C.byref(vb) if vb else None, C.byref(vy), 3) == 1
    ws = torch.empty(lib.dnnca_conv_workspace_bytes(9, cin, cout), dtype=torch.uint8, device='cuda')
    scratch = torch.zeros(lib.dnnca_conv2d_fold_scratch_bytes(cin, cout) // 4, dtype=torch.float32, device='cuda')
    stats = torch.zeros(2 * cout, dtype=torch.float64, device='cuda')
    code, alpha = (N.ACT_RELU, 0.0) if act == 'relu' else (N.ACT_LEAKY, 0.3)
    N.call('dnnca_conv2d_fprop_affine', None, C.byref(va), C.byref(vb) if vb else None, N.ptr(fa), N.ptr(fb), N.ptr(wd), N.ptr(bd),
           C.byref(vy), code, alpha, N.ptr(stats), N.ptr(ws), ws.numel(), N.ptr(scratch))
    torch.cuda.synchronize()
    got = y.float().cpu().numpy()
    ref = y_ref.detach().numpy()
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 2e-2 * scale, (np.abs(got - ref).max(), scale)
    assert rel_l2(got, ref) <= 8e-3, rel_l2(got, ref)
    # the border pixels are where the class bias matters: check them on their own
    border = np.zeros((h, w), bool)
    border[[0, -1], :] = True
    border[:, [0, -1]] = True
    assert rel_l2(got[:, border], ref[:, border]) <= 8e-3
    # BatchNorm statistics of the STORED output
    st = stats.cpu().numpy()
    np.testing.assert_allclose(st[:cout], got.astype(np.float64).sum((0, 1, 2)), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(st[cout:], (got.astype(np.float64) ** 2).sum((0, 1, 2)), rtol=1e-5, atol=1e-3)
    # wgrad through the fold
    dw = torch.zeros(3, 3, cin, cout, dtype=torch.float32, device='cuda')
    db = torch.zeros(cout, dtype=torch.float32, device='cuda')
    N.call('dnnca_conv2d_wgrad_affine', None, C.byref(va), C.byref(vb) if vb else None, N.ptr(fa), N.ptr(fb), C.byref(vdz), N.ptr(dw),
           N.ptr(db), N.ptr(scratch))
    torch.cuda.synchronize()
    assert rel_l2(dw.cpu().numpy(), gw.numpy()) <= 5e-3, rel_l2(dw.cpu().numpy(), gw.numpy())
    np.testing.assert_allclose(db.cpu().numpy(), dz.sum((0, 1, 2)).numpy(), rtol=1e-4, atol=1e-3)


def test_fold_is_refused_for_unsupported_shapes(N):
    lib = N.lib()
    x = torch.zeros(1, 8, 8, 32, dtype=torch.bfloat16, device='cuda')
    x2 = torch.zeros(1, 8, 8, 32, dtype=torch.bfloat16, device='cuda')
    y = torch.zeros(1, 8, 8, 32, dtype=torch.bfloat16, device='cuda')
    vx, vx2, vy = N.tensor_view(x), N.tensor_view(x2), N.tensor_view(y)
    assert lib.dnnca_conv2d_fold_supported(C.byref(vx), C.byref(vx2), C.byref(vy), 3) == 0      # [32+32]: two inputs need 64-multiples
    assert lib.dnnca_conv2d_fold_supported(C.byref(vx), None, C.byref(vy), 1) == 0              # 1x1
    xf = torch.zeros(1, 8, 8, 32, dtype=torch.float32, device='cuda')
    vxf = N.tensor_view(xf)
    assert lib.dnnca_conv2d_fold_supported(C.byref(vxf), None, C.byref(vy), 3) == 0             # fp32 mode
    ws = torch.empty(1 << 20, dtype=torch.uint8, device='cuda')
    sc = torch.zeros(26 * 32, dtype=torch.float32, device='cuda')
    w = torch.zeros(3, 3, 64, 32, device='cuda')
    with pytest.raises(N.DnncaError, match='fold'):
        N.call('dnnca_conv2d_fprop_affine', None, C.byref(vx), C.byref(vx2), None, None, N.ptr(w), None, C.byref(vy), 0, 0.0, None,
               N.ptr(ws), ws.numel(), N.ptr(sc))


@pytest.mark.parametrize('c,h,w', [(64, 16, 24), (32, 8, 8), (136, 6, 10), (12, 8, 8)])
def test_maxpool_with_folded_affine(N, c, h, w):
    rng = np.random.default_rng(c)
    n = 2
    a = bf(rng.normal(0, 1, (n, h, w, c)).astype(np.float32))
    aff = np.concatenate([rng.uniform(-1.5, 1.5, c), rng.normal(0, 0.3, c)]).astype(np.float32)
    xr = a * torch.tensor(aff[:c]) + torch.tensor(aff[c:])
    y_ref, idx_ref = ops.maxpool(xr, 2, return_indices=True)
    xa = a.to('cuda').bfloat16().contiguous()
    fa = torch.tensor(aff).cuda()
    y = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device='cuda')
    idx = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device='cuda')
    stats = torch.zeros(2 * c, dtype=torch.float64, device='cuda')
    va, vy = N.tensor_view(xa), N.tensor_view(y)
    N.call('dnnca_maxpool2x2_fwd_affine', None, C.byref(va), N.ptr(fa), C.byref(vy), N.ptr(idx), N.ptr(stats))
    torch.cuda.synchronize()
    got = y.float().cpu().numpy()
    np.testing.assert_array_equal(got, y_ref.bfloat16().float().numpy())
    # argmax: identical wherever the window's best element is unique in the affine image
    win = xr.numpy().reshape(n, h // 2, 2, w // 2, 2, c).transpose(0, 1, 3, 5, 2, 4).reshape(n, h // 2, w // 2, c, 4)
    s = np.sort(win, -1)
    clear = s[..., 3] > s[..., 2]
    assert np.array_equal(idx.cpu().numpy()[clear], idx_ref.numpy()[clear])
    st = stats.cpu().numpy()
    np.testing.assert_allclose(st[:c], got.astype(np.float64).sum((0, 1, 2)), rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(st[c:], (got.astype(np.float64) ** 2).sum((0, 1, 2)), rtol=1e-6, atol=1e-4)


OPTS = dict(n_filters_first=32, n_downsample=2, rate=2, kernel_size=3, conv_stride=1, bn=True, padding='same')


def _run(fold, weights, x, y, train=True):
    from dnncancerannotator_b200.models import tf_models
    os.environ['DNNCA_BN_FOLD'] = '1' if fold else '0'
    try:
        m = tf_models.UNetAnnotator(**OPTS, dtype='bf16')
        m.build((None, 64, 64, 3))
        m.compile(loss=dict(class_name='WeightedCrossentropy', config=dict(weight_mul=3.0)))
        m.set_weights(weights)
        outs = []
        for _ in range(4):                     # eager, eager, captured graph, replay
            per = m.forward_backward(x, y).cpu().numpy().copy()
            outs.append((per, m.last_logits.cpu().numpy().copy(), m.get_grads()))
        plan = m._plan(x.shape[0], 64, 64)
        ev = m(x).cpu().numpy().copy()
        return outs, plan, ev, m
    finally:
        os.environ.pop('DNNCA_BN_FOLD', None)


def test_folded_plan_matches_unfolded_plan_and_oracle():
    from dnncancerannotator_b200 import runtime as R
    from dnncancerannotator_b200.synthetic import make_slices
    ref = rm.build_model('UNetAnnotator', OPTS, (None, 64, 64, 3), seed=4)
    ref.randomize_bn(seed=6)
    x, y = make_slices(4, 64, 64, 3, seed=9)
    r = ref.train_step_grads(x, y, dict(weight_mul=3.0))
    fo, fplan, fev, fm = _run(True, ref.get_weights(), x, y)
    uo, uplan, uev, um = _run(False, ref.get_weights(), x, y)
    nf = sum(1 for op in fplan.ops if isinstance(op, R.BNOp) and op.folded)
    assert nf >= 6 and fplan.n_folded_bn == nf, nf               # enc bn0/bn1/pool_bn, dec tconv_bn/bn0 at the wide levels
    assert not any(isinstance(op, R.BNOp) and op.folded for op in uplan.ops)
    assert fplan.activation_bytes() < uplan.activation_bytes()   # the folded outputs are never allocated
    names = [k for k in ref.trainable if not k.endswith('/tconv/bias')]
    cat = lambda g: np.concatenate([np.asarray(g[k]).ravel() for k in names])
    allr = np.concatenate([r['grads'][k].numpy().ravel() for k in names])
    rl = r['logits'].numpy()
    for rep in range(4):                       # every execution mode agrees with itself
        assert rel_l2(fo[rep][1], fo[0][1]) < 1e-6
    f_log, u_log = rel_l2(fo[3][1], rl), rel_l2(uo[3][1], rl)
    f_g, u_g = rel_l2(cat(fo[3][2]), allr), rel_l2(cat(uo[3][2]), allr)
    # one bf16 rounding per BN less: the fold must not be further from the fp32 oracle than the unfolded plan (+ slack)
    assert f_log <= 1.2 * u_log + 2e-3, (f_log, u_log)
    assert f_g <= 1.2 * u_g + 5e-3, (f_g, u_g)
    assert abs(fo[3][0].mean() - r['data_loss']) <= 3e-3 * abs(r['data_loss'])
    # inference mode (moving statistics) folds too
    e = ref.forward(x, training=False)
    assert np.abs(fev - e['probs'].numpy()).max() <= 3e-2
    assert np.abs(fev - uev).max() <= 3e-2
    # BN moving statistics are updated exactly as before
    wf, wu = fm.get_weights(), um.get_weights()
    for k in wf:
        if 'moving' in k:
            np.testing.assert_allclose(wf[k], wu[k], rtol=2e-2, atol=2e-3)


# (n, h, w, cout of the layer = K of the dgrad GEMM, ca = channels of dx (the BN output), cb = channels of dx2)
BNR_SHAPES = [(2, 32, 24, 64, 64, 0), (1, 16, 40, 64, 128, 0), (2, 18, 22, 64, 64, 64), (1, 34, 16, 128, 128, 128),
              (2, 16, 16, 256, 256, 0), (2, 16, 16, 32, 32, 0), (2, 8, 8, 12, 12, 0)]


@pytest.mark.parametrize('shape', BNR_SHAPES)
def test_conv_dgrad_with_bn_backward_sums(N, shape):
    """dnnca_conv2d_dgrad_bnreduce == dnnca_conv2d_dgrad followed by dnnca_bn_bwd_reduce over (bn_x, dx): the gradient
    is bit-identical (same kernel arithmetic) and the sums agree to fp32 partial-sum round-off."""
    n, h, w, cout, ca, cb = shape
    rng = np.random.default_rng(abs(hash(shape)) % 2 ** 31)
    lib = N.lib()
    cin = ca + cb
    d = lambda t, dt=torch.bfloat16: torch.as_tensor(t).to('cuda').to(dt).contiguous()
    dz = d(rng.normal(0, 1.0, (n, h, w, cout)).astype(np.float32))
    wt = d((rng.normal(0, 1.0, (3, 3, cin, cout)) / np.sqrt(9 * cout)).astype(np.float32), torch.float32)
    bn_x = d(np.maximum(rng.normal(0.4, 1.0, (n, h, w, ca)), 0).astype(np.float32))
    mi = d(np.concatenate([rng.normal(0.5, 0.2, ca), rng.uniform(0.5, 2.0, ca)]).astype(np.float32), torch.float32)
    ws = torch.empty(lib.dnnca_conv_workspace_bytes(9, cin, cout), dtype=torch.uint8, device='cuda')
    outs = []
    for fused in (False, True):
        dx = torch.full((n, h, w, ca), 7.0, dtype=torch.bfloat16, device='cuda')
        dx2 = torch.full((n, h, w, cb), 7.0, dtype=torch.bfloat16, device='cuda') if cb else None
        sums = torch.zeros(2 * ca, dtype=torch.float64, device='cuda')
        vdz, vdx, vx = N.tensor_view(dz), N.tensor_view(dx), N.tensor_view(bn_x)
        vdx2 = N.tensor_view(dx2) if cb else None
        if fused:
            N.call('dnnca_conv2d_dgrad_bnreduce', None, C.byref(vdz), N.ptr(wt), C.byref(vdx), C.byref(vdx2) if cb else None, 3,
                   C.byref(vx), N.ptr(mi), N.ptr(sums), N.ptr(ws), ws.numel())
        else:
            N.call('dnnca_conv2d_dgrad', None, C.byref(vdz), N.ptr(wt), C.byref(vdx), C.byref(vdx2) if cb else None, 3, None,
                   N.ACT_NONE, 0.0, N.ptr(ws), ws.numel())
            N.call('dnnca_bn_bwd_reduce', None, C.byref(vx), C.byref(vdx), N.ptr(mi), N.ptr(sums))
        torch.cuda.synchronize()
        outs.append((dx.float().cpu().numpy(), dx2.float().cpu().numpy() if cb else None, sums.cpu().numpy()))
    (dx0, dxb0, s0), (dx1, dxb1, s1) = outs
    np.testing.assert_array_equal(dx0, dx1)
    if cb:
        np.testing.assert_array_equal(dxb0, dxb1)
    # independent fp64 sums from the stored gradient
    xh = (bn_x.float().cpu().numpy().astype(np.float64) - mi[:ca].cpu().numpy()) * mi[ca:].cpu().numpy()
    want = np.concatenate([dx0.astype(np.float64).sum((0, 1, 2)), (dx0.astype(np.float64) * xh).sum((0, 1, 2))])
    scale = np.abs(want).max()
    np.testing.assert_allclose(s1, want, rtol=2e-4, atol=2e-5 * scale)
    np.testing.assert_allclose(s0, want, rtol=2e-4, atol=2e-5 * scale)


@pytest.mark.parametrize('shape', [(2, 8, 12, 64, 128), (1, 16, 16, 128, 256), (2, 8, 8, 12, 12)])
def test_tconv_dgrad_with_bn_backward_sums(N, shape):
    n, h, w, cout, cin = shape                  # ConvT cin -> cout, x [n,h,w,cin], dy [n,2h,2w,cout]
    rng = np.random.default_rng(abs(hash(shape)) % 2 ** 31)
    lib = N.lib()
    d = lambda t, dt=torch.bfloat16: torch.as_tensor(t).to('cuda').to(dt).contiguous()
    dy = d(rng.normal(0, 1.0, (n, 2 * h, 2 * w, cout)).astype(np.float32))
    kt = d((rng.normal(0, 1.0, (2, 2, cout, cin)) / np.sqrt(4 * cout)).astype(np.float32), torch.float32)
    bn_x = d(np.maximum(rng.normal(0.4, 1.0, (n, h, w, cin)), 0).astype(np.float32))
    mi = d(np.concatenate([rng.normal(0.5, 0.2, cin), rng.uniform(0.5, 2.0, cin)]).astype(np.float32), torch.float32)
    ws = torch.empty(lib.dnnca_conv_workspace_bytes(4, cin, cout), dtype=torch.uint8, device='cuda')
    outs = []
    for fused in (False, True):
        dx = torch.full((n, h, w, cin), 7.0, dtype=torch.bfloat16, device='cuda')
        sums = torch.zeros(2 * cin, dtype=torch.float64, device='cuda')
        vdy, vdx, vx = N.tensor_view(dy), N.tensor_view(dx), N.tensor_view(bn_x)
        if fused:
            N.call('dnnca_convtranspose2x2_dgrad_bnreduce', None, C.byref(vdy), N.ptr(kt), C.byref(vdx), C.byref(vx), N.ptr(mi),
                   N.ptr(sums), N.ptr(ws), ws.numel())
        else:
            N.call('dnnca_convtranspose2x2_dgrad', None, C.byref(vdy), N.ptr(kt), C.byref(vdx), None, N.ACT_NONE, 0.0, N.ptr(ws),
                   ws.numel())
            N.call('dnnca_bn_bwd_reduce', None, C.byref(vx), C.byref(vdx), N.ptr(mi), N.ptr(sums))
        torch.cuda.synchronize()
        outs.append((dx.float().cpu().numpy(), sums.cpu().numpy()))
    (dx0, s0), (dx1, s1) = outs
    np.testing.assert_array_equal(dx0, dx1)
    xh = (bn_x.float().cpu().numpy().astype(np.float64) - mi[:cin].cpu().numpy()) * mi[cin:].cpu().numpy()
    want = np.concatenate([dx0.astype(np.float64).sum((0, 1, 2)), (dx0.astype(np.float64) * xh).sum((0, 1, 2))])
    scale = np.abs(want).max()
    np.testing.assert_allclose(s1, want, rtol=2e-4, atol=2e-5 * scale)


def test_plan_fuses_bn_backward_reductions():
    """Every BatchNorm whose output has one reader loses its bn_bwd_reduce pass; gradients agree with the unfused plan."""
    from dnncancerannotator_b200 import runtime as R
    from dnncancerannotator_b200.models import tf_models
    from dnncancerannotator_b200.synthetic import make_slices
    ref = rm.build_model('UNetAnnotator', OPTS, (None, 64, 64, 3), seed=4)
    ref.randomize_bn(seed=6)
    x, y = make_slices(4, 64, 64, 3, seed=9)
    res = {}
    for fuse in ('1', '0'):
        os.environ['DNNCA_BN_REDUCE_FUSE'] = fuse
        try:
            m = tf_models.UNetAnnotator(**OPTS, dtype='bf16')
            m.build((None, 64, 64, 3))
            m.compile(loss=dict(class_name='WeightedCrossentropy', config=dict(weight_mul=3.0)))
            m.set_weights(ref.get_weights())
            m.forward_backward(x, y)
            plan = m._plan(4, 64, 64)
            res[fuse] = (m.get_grads(), sum(1 for op in plan.ops if isinstance(op, R.BNOp) and op.reduce_fused),
                         sum(1 for op in plan.ops if isinstance(op, R.BNOp)))
        finally:
            os.environ.pop('DNNCA_BN_REDUCE_FUSE', None)
    (g1, n1, nb), (g0, n0, _) = res['1'], res['0']      # (the fusion is opt-in: DNNCA_BN_REDUCE_FUSE=1)
    # n_downsample=2: enc 2 x (bn0, pool_bn) + dec 2 x (tconv_bn, bn0) + dec u0 bn1 (read by the next ConvT); skips and the last BN are not
    assert n0 == 0 and n1 >= 8 and n1 < nb, (n1, nb)
    names = [k for k in g1 if not k.endswith('/tconv/bias')]
    a = np.concatenate([g1[k].ravel() for k in names])
    b = np.concatenate([g0[k].ravel() for k in names])
    assert rel_l2(a, b) <= 2e-3, rel_l2(a, b)
