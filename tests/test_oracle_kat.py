"""First-principles known-answer tests that pin the oracle (SURVEY.md 8c "pins").

The reference holds no golden vectors for this path, so these KATs are derived
from the defining formulas of the Keras ops the reference calls.
"""
import math

import numpy as np
import torch

from oracle import ref_ops as ops
from oracle import ref_numpy as rn


def test_conv_delta_image_places_flipped_kernel():
    # cross-correlation: a delta at (3,4) produces k[a,c] at (3-a+1, 4-c+1)
    k = np.arange(1, 10, dtype=np.float32).reshape(3, 3, 1, 1)
    x = np.zeros((1, 8, 8, 1), np.float32)
    x[0, 3, 4, 0] = 1.0
    y = ops.conv2d(torch.tensor(x), torch.tensor(k)).numpy()[0, :, :, 0]
    for a in range(3):
        for c in range(3):
            assert y[3 - a + 1, 4 - c + 1] == k[a, c, 0, 0]
    assert y.sum() == k.sum()


def test_conv_same_zero_padding_corner():
    x = np.ones((1, 4, 4, 2), np.float32)
    k = np.ones((3, 3, 2, 3), np.float32)
    y = ops.conv2d(torch.tensor(x), torch.tensor(k), torch.tensor([0.5, 0., -1.])).numpy()
    assert y[0, 0, 0, 0] == 4 * 2 + 0.5      # corner sees 2x2 taps
    assert y[0, 0, 1, 1] == 6 * 2            # edge sees 2x3
    assert y[0, 1, 1, 2] == 9 * 2 - 1        # interior


def test_tconv_block_pattern():
    # k=s=2: each input pixel paints a 2x2 block, out[2i+a,2j+b,co] = sum_ci x*K[a,b,co,ci]
    x = np.zeros((1, 2, 2, 2), np.float32)
    x[0, 1, 0] = [1.0, 2.0]
    K = np.arange(2 * 2 * 3 * 2, dtype=np.float32).reshape(2, 2, 3, 2)
    b = np.array([10., 20., 30.], np.float32)
    y = ops.conv2d_transpose(torch.tensor(x), torch.tensor(K), torch.tensor(b)).numpy()
    for a in range(2):
        for c in range(2):
            np.testing.assert_allclose(y[0, 2 + a, 0 + c], K[a, c] @ x[0, 1, 0] + b)
    np.testing.assert_allclose(y[0, 0, 0], b)


def test_bn_constant_channel_gives_beta_and_moving_update():
    x = torch.full((2, 4, 4, 3), 7.0)
    x[..., 1] = torch.arange(32, dtype=torch.float32).reshape(2, 4, 4)
    g, b = torch.tensor([2., 3., 4.]), torch.tensor([.1, .2, .3])
    y, mm, mv = ops.batchnorm(x, g, b, torch.zeros(3), torch.ones(3), training=True)
    np.testing.assert_allclose(y[..., 0].numpy(), 0.1, atol=1e-6)
    np.testing.assert_allclose(y[..., 2].numpy(), 0.3, atol=1e-6)
    # moving stats: momentum .99, unbiased variance for the moving var
    np.testing.assert_allclose(mm.numpy(), [0.07, 0.155, 0.07], rtol=1e-6)
    var_unb = np.var(np.arange(32.0), ddof=1)
    np.testing.assert_allclose(mv.numpy(), [0.99, 0.99 + 0.01 * var_unb, 0.99], rtol=1e-6)
    # biased variance normalises: channel 1 has unit (biased) variance up to eps
    xh = (y[..., 1] - 0.2) / 3.0
    np.testing.assert_allclose(float((xh ** 2).mean()), np.var(np.arange(32.0)) / (np.var(np.arange(32.0)) + 1e-3), rtol=1e-5)


def test_bn_inference_uses_moving_stats():
    x = torch.randn(2, 3, 3, 2)
    y, _, _ = ops.batchnorm(x, None, torch.tensor([1., 2.]), torch.tensor([.5, -.5]), torch.tensor([4., 9.]), False)
    ref = (x - torch.tensor([.5, -.5])) / torch.sqrt(torch.tensor([4., 9.]) + 1e-3) + torch.tensor([1., 2.])
    np.testing.assert_allclose(y.numpy(), ref.numpy(), rtol=1e-6, atol=1e-6)


def test_pool_first_max_wins():
    y, idx = ops.maxpool(torch.zeros(1, 4, 4, 1), 2, return_indices=True)
    assert idx.flatten().tolist() == [0, 0, 0, 0]
    x = torch.tensor([[1., 2.], [2., 0.]]).reshape(1, 2, 2, 1)
    y, idx = ops.maxpool(x, 2, return_indices=True)
    assert float(y) == 2.0 and int(idx) == 1
    # gradient goes to that first maximum only
    x = x.clone().requires_grad_(True)
    ops.maxpool(x, 2).sum().backward()
    assert x.grad.flatten().tolist() == [0., 1., 0., 0.]
    yn, idxn = rn.maxpool2x2_fwd(x.detach().numpy())
    assert int(idxn.item()) == 1


def test_loss_all_zero_labels_is_plain_bce_mean():
    z = torch.randn(3, 8, 8, 1)
    y = torch.zeros(3, 8, 8)
    got = ops.weighted_crossentropy(y, z, weight_mul=3.0)
    ref = torch.nn.functional.softplus(z[..., 0]).mean(dim=(1, 2))   # BCE(y=0) = softplus(z)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-6)


def test_loss_weight_from_positive_rate():
    # r = 1/64 -> w = 3*64 = 192 ; with z = 0 every pixel costs log 2
    y = torch.zeros(1, 8, 8)
    y[0, 2, 5] = 1.0
    z = torch.zeros(1, 8, 8, 1)
    got = float(ops.weighted_crossentropy(y, z, weight_mul=3.0)[0])
    assert math.isclose(got, math.log(2) * (63 + 192) / 64, rel_tol=1e-6)
    # explicit weight overrides the positive-rate branch (losses.py:25)
    got = float(ops.weighted_crossentropy(y, z, weight=5.0, weight_add=1.0, weight_mul=2.0)[0])
    assert math.isclose(got, math.log(2) * (63 + 11) / 64, rel_tol=1e-6)
    # empty batch (losses.py:22-23)
    assert ops.weighted_crossentropy(torch.zeros(0, 8, 8), torch.zeros(0, 8, 8, 1)).shape == (0,)


def test_bce_is_stable_for_large_logits():
    y = torch.tensor([[[1.0, 0.0]]])
    z = torch.tensor([[[[-200.0], [200.0]]]])
    got = ops.weighted_crossentropy(y, z, weight=1.0)
    assert math.isclose(float(got[0]), 200.0, rel_tol=1e-6)


def test_adam_keras_form_first_step():
    p, m, v = ops.adam_step(torch.tensor([1.0]), torch.tensor([1.0]), torch.zeros(1), torch.zeros(1), step=1)
    lr_t = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
    expect = 1.0 - lr_t * 0.1 / (math.sqrt(0.001) + 1e-7)
    assert math.isclose(float(p), expect, rel_tol=1e-7)
    assert math.isclose(float(m), 0.1, rel_tol=1e-6) and math.isclose(float(v), 0.001, rel_tol=1e-6)


def test_lr_schedule_matches_deploy_options_lambda():
    assert ops.lr_schedule(0) == 0.001
    assert ops.lr_schedule(999) == 0.001
    assert math.isclose(ops.lr_schedule(1000), 0.00096)
    assert math.isclose(ops.lr_schedule(2500), 0.001 * 0.96 ** 2)
