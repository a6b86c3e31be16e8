"""torch-CPU functional restatement of the Keras/TF ops on the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``); parity unpinned.

All tensors are NHWC like the reference's (``data.py:193-206`` produces
``[B,H,W,C]``); weights use the TF layouts.  Functions take and return torch
tensors so that ``torch.autograd`` can play the role of ``tf.GradientTape``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# layout helpers
# ----------------------------------------------------------------------------
def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


# ----------------------------------------------------------------------------
# activations  (components.py:323-335 ``solve_activation``)
# ----------------------------------------------------------------------------
def activation(x, act):
    """``act`` is None | 'relu' | 'sigmoid' | ('leaky', alpha).

    'relu' -> ``tf.keras.activations.relu`` ; the dict form in
    ``configs/additionals/leakyReLU.yaml:1-4`` resolves to
    ``tf.keras.layers.LeakyReLU(alpha=0.3)`` (components.py:329-333).
    """
    if act is None or act == 'linear':
        return x
    if act == 'relu':
        return torch.relu(x)
    if act == 'sigmoid':
        return torch.sigmoid(x)
    if isinstance(act, tuple) and act[0] == 'leaky':
        return F.leaky_relu(x, negative_slope=float(act[1]))
    raise ValueError(f'unknown activation {act!r}')


# ----------------------------------------------------------------------------
# Conv2D  (components.py:47-50,123-126 ; unet.py:241-244 ; multiresunet.py:51-52)
# ----------------------------------------------------------------------------
def conv2d(x, kernel, bias=None, padding='same', stride=1):
    """[TF-semantics] ``layers.Conv2D``: cross-correlation (no kernel flip),
    kernel ``[kh,kw,Cin,Cout]``; 'same' with odd k and stride 1 pads (k-1)/2
    zeros on each side; 'valid' pads nothing.
    """
    kh, kw, cin, cout = kernel.shape
    w = kernel.permute(3, 2, 0, 1)  # HWIO -> OIHW
    if padding == 'same':
        assert stride == 1 and kh % 2 == 1 and kw % 2 == 1, 'only odd k / stride 1 restated'
        pad = (kh // 2, kw // 2)
    elif padding == 'valid':
        pad = (0, 0)
    else:
        raise ValueError(padding)
    y = F.conv2d(_nchw(x), w, bias, stride=stride, padding=pad)
    return _nhwc(y)


# ----------------------------------------------------------------------------
# Conv2DTranspose k = s = rate (components.py:118-120 ; multiresunet.py:200-215)
# ----------------------------------------------------------------------------
def conv2d_transpose(x, kernel, bias=None, stride=2):
    """[TF-semantics] ``layers.Conv2DTranspose(k=s)``: kernel ``[kh,kw,Cout,Cin]``;
    with k == s both 'same' and 'valid' give an exactly s-times larger output and
    non-overlapping taps: ``out[s*i+a, s*j+b, co] = sum_ci x[i,j,ci]*K[a,b,co,ci] + b[co]``.
    """
    kh, kw, cout, cin = kernel.shape
    assert kh == stride and kw == stride, 'only k == s restated (the only use in the reference)'
    w = kernel.permute(3, 2, 0, 1)  # -> torch conv_transpose layout [Cin,Cout,kh,kw]
    y = F.conv_transpose2d(_nchw(x), w, bias, stride=stride)
    return _nhwc(y)


# ----------------------------------------------------------------------------
# MaxPool2D (components.py:54 ; multiresunet.py:183-195)
# ----------------------------------------------------------------------------
def maxpool(x, rate=2, return_indices=False):
    """[TF-semantics] ``MaxPool2D([r,r], strides=r)``, VALID.  Gradient goes to
    the first maximum in row-major window order (TF CPU and torch CPU agree).

    With ``return_indices`` also returns the window-local argmax (0..r*r-1,
    row-major) as uint8 ``[B,H/r,W/r,C]`` -- the layout the CUDA path stores.
    """
    y, idx = F.max_pool2d(_nchw(x), rate, rate, return_indices=True)
    out = _nhwc(y)
    if not return_indices:
        return out
    h, w = x.shape[1], x.shape[2]
    iy = idx // w
    ix = idx % w
    local = (iy % rate) * rate + (ix % rate)
    return out, _nhwc(local).to(torch.uint8)


# ----------------------------------------------------------------------------
# BatchNormalization (components.py:57,59,130,131 ; multiresunet.py:53,120,124,150,162)
# ----------------------------------------------------------------------------
BN_MOMENTUM = 0.99   # [TF-semantics] keras default
BN_EPSILON = 1e-3    # [TF-semantics] keras default


def batchnorm(x, gamma, beta, moving_mean, moving_var, training,
              momentum=BN_MOMENTUM, eps=BN_EPSILON):
    """[TF-semantics] axis=-1 batch norm.

    training: normalise with the *biased* batch variance over (N,H,W); the new
    moving statistics are ``m*momentum + batch*(1-momentum)`` where the moving
    variance uses the *unbiased* batch variance (fused kernel behaviour).
    inference: moving statistics.  ``gamma`` may be None (``scale=False``,
    multiresunet.py:53).

    Returns (y, new_moving_mean, new_moving_var).
    """
    if training:
        n = x.shape[0] * x.shape[1] * x.shape[2]
        mean = x.mean(dim=(0, 1, 2))
        var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
        xhat = (x - mean) * torch.rsqrt(var + eps)
        with torch.no_grad():
            unbiased = var * (n / max(n - 1, 1))
            new_mm = moving_mean * momentum + mean * (1 - momentum)
            new_mv = moving_var * momentum + unbiased * (1 - momentum)
    else:
        xhat = (x - moving_mean) * torch.rsqrt(moving_var + eps)
        new_mm, new_mv = moving_mean, moving_var
    y = xhat if gamma is None else xhat * gamma
    y = y + beta
    return y, new_mm, new_mv


# ----------------------------------------------------------------------------
# Loss (losses.py:17-37, 60-72, 87-102)
# ----------------------------------------------------------------------------
def positive_rate(label):
    """losses.py:87-102 ``tf_get_positive_rate``: sum(label)/numel over the whole
    (per-replica) batch; asserts 0 <= label <= 1."""
    assert float(label.max()) <= 1.0 and float(label.min()) >= 0.0
    return label.sum() / label.numel()


def gaussian_filter2d(label, filter_size=6, sigma=3.0):
    """losses.py:62-67 label smoothing. [TF-semantics] ``tfa.image.gaussian_filter2d``
    with the default ``padding='REFLECT'``: separable normalised gaussian of
    ``filter_size`` taps centred at (size-1)/2 ... for an even size TFA pads
    (size-1)//2 before and size - 1 - (size-1)//2 after."""
    # TFA builds the kernel on range(-size//2 + 1, size//2 + 1) (6 -> -2..3)
    k = torch.arange(-filter_size // 2 + 1, filter_size // 2 + 1, dtype=label.dtype, device=label.device)
    g = torch.exp(-(k ** 2) / (2.0 * sigma ** 2))
    g = g / g.sum()
    pad_before = (filter_size - 1) // 2
    pad_after = filter_size - 1 - pad_before
    x = label.unsqueeze(1)  # [B,1,H,W]
    x = F.pad(x, (pad_before, pad_after, pad_before, pad_after), mode='reflect')
    x = F.conv2d(x, g.view(1, 1, 1, -1))
    x = F.conv2d(x, g.view(1, 1, -1, 1))
    return x.squeeze(1)


def weighted_crossentropy(label, logits, weight=None, weight_add=0.0, weight_mul=1.0):
    """losses.py:17-37 ``tf_weighted_crossentropy`` with ``from_logits=True``
    (the only way ``TFWeightedCrossentropy.call`` invokes it, losses.py:68-71).

    label ``[B,H,W]`` in [0,1]; logits ``[B,H,W,1]``.  Returns the per-sample loss
    ``[B]``.  [TF-semantics] BCE from logits = max(z,0) - z*y + log1p(exp(-|z|)),
    mean over the (size-1) channel axis, times the sample-weight mask, then
    ``reduce_mean`` over H,W (losses.py:36).
    """
    if label.shape[0] == 0:  # losses.py:22-23
        return logits.new_zeros([0])
    if weight is None:  # losses.py:25-27
        r = positive_rate(label)
        weight = 1.0 / r if float(r) > 0.0 else torch.tensor(1.0, dtype=label.dtype, device=label.device)
    weight = weight_mul * weight + weight_add  # losses.py:29
    assert float(weight) >= 0.0  # losses.py:30
    mask = label * (weight - 1) + torch.ones_like(label)  # losses.py:31
    z = logits[..., 0]
    # [TF-semantics] tf.nn.sigmoid_cross_entropy_with_logits builds max(z,0) and -|z| with `where(z >= 0, ...)`, so its
    # autodiff gives sigmoid(z) - y everywhere INCLUDING z == 0 (clamp/abs would give 1 - y there); dead-ReLU pixels of
    # a zero-bias head sit at z == 0 exactly
    pos = z >= 0
    relu_z = torch.where(pos, z, torch.zeros_like(z))
    neg_abs_z = torch.where(pos, -z, z)
    bce = relu_z - z * label + torch.log1p(torch.exp(neg_abs_z))
    loss = bce * mask
    return loss.mean(dim=(1, 2))


def l2_regularizer(kernels, l2):
    """kernel_regularizer.yaml:1-4 -> [TF-semantics] ``keras.regularizers.L2``:
    ``l2 * sum(w**2)`` per regularised kernel, summed into the training loss."""
    return sum(l2 * (k ** 2).sum() for k in kernels)


# ----------------------------------------------------------------------------
# Adam (engine.py:276-284)  [TF-semantics] keras OptimizerV2 Adam, non-amsgrad
# ----------------------------------------------------------------------------
def adam_step(param, grad, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
    """``step`` is the 1-based iteration count *after* increment.

    m <- b1*m + (1-b1)*g ; v <- b2*v + (1-b2)*g^2 ;
    p <- p - lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)   (eps outside the
    bias correction, unlike ``torch.optim.Adam``).
    """
    m = beta1 * m + (1 - beta1) * grad
    v = beta2 * v + (1 - beta2) * grad * grad
    lr_t = lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    param = param - lr_t * m / (torch.sqrt(v) + eps)
    return param, m, v


def lr_schedule(step, current_lr=None):
    """deploy_options.yaml:3 ``lambda epoch, current_lr: 0.001 * 0.96 ** (epoch // 1000)``
    evaluated with the 0-based step index as "epoch" (engine.py:98-100,126-135)."""
    return 0.001 * 0.96 ** (step // 1000)
