"""bf16-storage emulation of the oracle (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

The CUDA path stores every activation and activation-gradient in bf16 and feeds the tensor
cores bf16 copies of the fp32 master weights.  ``emulate_bf16()`` patches ``ref_ops`` so that
the torch-CPU restatement rounds at the same points (conv/ConvT outputs after the activation,
BatchNorm outputs, the network input, the weights of layers wide enough for tcgen05, and the
matching gradients).  It answers the question "how far is ANY bf16-storage implementation of
this model from the fp32 reference?", which for the BatchNorm configs at random initialisation
is far (DESIGN.md "bf16 sensitivity"): the CUDA path is then held to "no worse than the
emulation" instead of an absolute bound it cannot meet.
"""
import contextlib

import torch

from . import ref_ops as ops


_MANTISSA = 7     # stored mantissa bits of the emulated format: 7 = bfloat16, 10 = TF32 / fp16, 23 = fp32 (no rounding)


def round_mantissa(x, bits=None):
    """Round-to-nearest-even of a float32 tensor to ``bits`` stored mantissa bits (exponent range kept): bits=7 is
    exactly ``x.bfloat16().float()``; bits=10 is what a TF32 tensor-core multiply sees of its operands."""
    bits = _MANTISSA if bits is None else bits
    if bits >= 23:
        return x
    if bits == 7:
        return x.bfloat16().float()
    drop = 23 - bits
    i = x.detach().contiguous().view(torch.int32)
    lsb = (i >> drop) & 1
    i = (i + ((1 << (drop - 1)) - 1) + lsb) & ~((1 << drop) - 1)
    return i.view(torch.float32)


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return round_mantissa(x)

    @staticmethod
    def backward(ctx, g):
        return round_mantissa(g)


def _rw(k):
    """low-precision copy of a weight with a straight-through gradient to the fp32 master."""
    return k + (round_mantissa(k.detach()) - k.detach())


@contextlib.contextmanager
def emulate_bf16(round_weights_min_channels=16, mantissa_bits=7, round_conv_outputs=False):
    global _MANTISSA
    saved_bits, _MANTISSA = _MANTISSA, mantissa_bits
    oc, ot, ob, oa = ops.conv2d, ops.conv2d_transpose, ops.batchnorm, ops.activation
    r = _Round.apply

    def conv2d(x, k, b=None, padding='same', stride=1):
        tc = k.shape[2] >= round_weights_min_channels and k.shape[3] >= round_weights_min_channels
        y = oc(x, _rw(k) if tc else k, b, padding, stride)
        # MultiResUnet's conv2d_bn is Conv2D -> BN -> activation: the conv output itself is a stored tensor there
        return r(y) if round_conv_outputs else y

    def activation(x, act):
        return r(oa(x, act))

    def conv2d_transpose(x, k, b=None, stride=2):
        tc = k.shape[2] >= round_weights_min_channels and k.shape[3] >= round_weights_min_channels
        return r(ot(x, _rw(k) if tc else k, b, stride))

    def batchnorm(*a, **kw):
        y, mm, mv = ob(*a, **kw)
        return r(y), mm, mv

    ops.conv2d, ops.conv2d_transpose, ops.batchnorm, ops.activation = conv2d, conv2d_transpose, batchnorm, activation
    try:
        yield
    finally:
        ops.conv2d, ops.conv2d_transpose, ops.batchnorm, ops.activation = oc, ot, ob, oa
        _MANTISSA = saved_bits
