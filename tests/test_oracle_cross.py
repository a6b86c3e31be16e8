"""The two independent restatements (torch/autograd vs loop-level numpy fp64)
must agree on tiny shapes -- forward and backward."""
import numpy as np
import pytest
import torch

from oracle import ref_ops as ops
from oracle import ref_numpy as rn

T = lambda a: torch.tensor(a, dtype=torch.float64)


@pytest.mark.parametrize('k', [1, 3])
def test_conv_fwd_bwd(k):
    rng = np.random.default_rng(0)
    x, w, b = rng.normal(size=(2, 6, 5, 3)), rng.normal(size=(k, k, 3, 4)), rng.normal(size=4)
    dy = rng.normal(size=(2, 6, 5, 4))
    xt, wt, bt = T(x).requires_grad_(), T(w).requires_grad_(), T(b).requires_grad_()
    y = ops.conv2d(xt, wt, bt)
    y.backward(T(dy))
    np.testing.assert_allclose(y.detach().numpy(), rn.conv2d_same_fwd(x, w, b), rtol=1e-10, atol=1e-10)
    dx, dk, db = rn.conv2d_same_bwd(x, w, dy)
    np.testing.assert_allclose(xt.grad.numpy(), dx, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(wt.grad.numpy(), dk, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(bt.grad.numpy(), db, rtol=1e-10, atol=1e-10)


def test_tconv_fwd_bwd():
    rng = np.random.default_rng(1)
    x, w, b = rng.normal(size=(2, 3, 4, 5)), rng.normal(size=(2, 2, 3, 5)), rng.normal(size=3)
    dy = rng.normal(size=(2, 6, 8, 3))
    xt, wt, bt = T(x).requires_grad_(), T(w).requires_grad_(), T(b).requires_grad_()
    y = ops.conv2d_transpose(xt, wt, bt)
    y.backward(T(dy))
    np.testing.assert_allclose(y.detach().numpy(), rn.tconv2x2_fwd(x, w, b), rtol=1e-10, atol=1e-10)
    dx, dk, db = rn.tconv2x2_bwd(x, w, dy)
    np.testing.assert_allclose(xt.grad.numpy(), dx, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(wt.grad.numpy(), dk, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(bt.grad.numpy(), db, rtol=1e-10, atol=1e-10)


def test_pool_fwd_bwd_with_ties():
    rng = np.random.default_rng(2)
    x = rng.integers(0, 3, size=(2, 6, 8, 3)).astype(np.float64)   # many ties
    dy = rng.normal(size=(2, 3, 4, 3))
    xt = T(x).requires_grad_()
    y, idx = ops.maxpool(xt, 2, return_indices=True)
    y.backward(T(dy))
    yn, idxn = rn.maxpool2x2_fwd(x)
    np.testing.assert_array_equal(y.detach().numpy(), yn)
    np.testing.assert_array_equal(idx.numpy(), idxn)
    np.testing.assert_array_equal(xt.grad.numpy(), rn.maxpool2x2_bwd(dy, idxn))


@pytest.mark.parametrize('scale', [True, False])
def test_bn_fwd_bwd(scale):
    rng = np.random.default_rng(3)
    x, dy = rng.normal(size=(2, 4, 4, 3)), rng.normal(size=(2, 4, 4, 3))
    gamma = rng.uniform(.5, 1.5, 3) if scale else None
    beta = rng.normal(size=3)
    xt = T(x).requires_grad_()
    gt = T(gamma).requires_grad_() if scale else None
    bt = T(beta).requires_grad_()
    y, _, _ = ops.batchnorm(xt, gt, bt, torch.zeros(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64), True)
    y.backward(T(dy))
    yn, mean, var, invstd = rn.bn_train_fwd(x, gamma, beta)
    np.testing.assert_allclose(y.detach().numpy(), yn, rtol=1e-10, atol=1e-10)
    dx, dg, db = rn.bn_train_bwd(x, gamma, dy, mean, invstd)
    np.testing.assert_allclose(xt.grad.numpy(), dx, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(bt.grad.numpy(), db, rtol=1e-10, atol=1e-10)
    if scale:
        np.testing.assert_allclose(gt.grad.numpy(), dg, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize('healthy', [False, True])
def test_loss_fwd_bwd(healthy):
    rng = np.random.default_rng(4)
    y = (rng.uniform(size=(3, 8, 8)) < (0.0 if healthy else 0.1)).astype(np.float64)
    z = rng.normal(size=(3, 8, 8, 1)) * 3
    zt = T(z).requires_grad_()
    per = ops.weighted_crossentropy(T(y), zt, weight_mul=3.0)
    per.mean().backward()
    pn, dz = rn.weighted_bce_fwd_bwd(y, z, weight_mul=3.0)
    np.testing.assert_allclose(per.detach().numpy(), pn, rtol=1e-10)
    np.testing.assert_allclose(zt.grad.numpy()[..., 0], dz, rtol=1e-9, atol=1e-14)


def test_leaky_relu_grad_from_output_sign():
    rng = np.random.default_rng(5)
    x, g = rng.normal(size=(50,)), rng.normal(size=(50,))
    xt = T(x).requires_grad_()
    y = ops.activation(xt, ('leaky', 0.3))
    y.backward(T(g))
    np.testing.assert_allclose(xt.grad.numpy(), rn.act_bwd(y.detach().numpy(), g, ('leaky', 0.3)))
