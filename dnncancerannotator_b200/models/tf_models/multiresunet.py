"""MultiResUnet with the reference's signature (``multiresunet.py:167-223``),
forward / inference only (BASELINE configs[4] is a forward sweep; SURVEY.md 2 "M8").

Inference-time lowering: ``conv2d_bn`` = Conv2D(no bias) -> BN(scale=False) -> act
(multiresunet.py:31-60) with moving statistics is a per-output-channel affine of the conv
result, i.e. conv with scaled kernels + a bias (exact: zero padding precedes the conv);
the block tails BN -> add -> relu -> BN (multiresunet.py:119-124, 148-150) run as one
``dnnca_add_relu_affine`` pass; the three chain convs of a MultiRes block write disjoint
channel ranges of one buffer (``concatenate``, multiresunet.py:119).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ... import native as N
from ... import runtime as R
from ...keras_like import Model
from .components import glorot_uniform


class _FoldedConv(R.Op):
    """conv2d_bn at inference: y = act(conv(x, K*s) + t), s/t from the BN moving statistics.
    The folded kernel/bias are recomputed from the current variables at every forward
    (a C-sized device op on fp32 masters) so weight updates / loads are always honoured."""

    def __init__(self, plan, x, y, conv, bn, ksize, act):
        self.p, self.x, self.y, self.conv, self.bn, self.k, self.act = plan, x, y, conv, bn, ksize, act
        self.wf = self.bf = None

    def allocate(self, training):
        if self.wf is None:
            k = self.p.params.view(f'{self.conv}/kernel')
            self.wf = torch.empty_like(k)
            self.bf = torch.empty(k.shape[-1], dtype=torch.float32, device=self.p.device)

    def fwd(self, train):
        assert not train, 'MultiResUnet is forward/inference-only in this build'
        ps = self.p.params
        # host-side folding glue on parameter-sized tensors (not activation arithmetic)
        s = torch.rsqrt(ps.view(f'{self.bn}/moving_var') + R.BN_EPSILON)
        torch.mul(ps.view(f'{self.conv}/kernel'), s, out=self.wf)
        torch.addcmul(ps.view(f'{self.bn}/beta'), ps.view(f'{self.bn}/moving_mean'), s, value=-1.0, out=self.bf)
        N.call('dnnca_conv2d_fprop', N.stream_ptr(), self.x.ct(), None, N.ptr(self.wf), N.ptr(self.bf), self.y.ct(),
               self.k, self.act, 0.0, None, None, 0)


class _Affine:
    """scale|shift [2C] of an inference BatchNormalization, refreshed before use."""

    def __init__(self, plan, bn, c, scale=True):
        self.p, self.bn, self.c, self.scale = plan, bn, c, scale
        self.buf = None

    def __call__(self):
        ps = self.p.params
        if self.buf is None:
            self.buf = torch.empty(2 * self.c, dtype=torch.float32, device=self.p.device)
        N.call('dnnca_bn_inference_params', N.stream_ptr(), self.c,
               ps.ptr(f'{self.bn}/gamma') if self.scale else None, ps.ptr(f'{self.bn}/beta'), R.BN_EPSILON,
               ps.ptr(f'{self.bn}/moving_mean'), ps.ptr(f'{self.bn}/moving_var'), N.ptr(self.buf))
        return self.buf


class _Builder:
    """Walks the reference graph once to create variables (plan=None) or to emit ops."""

    def __init__(self, model, plan=None):
        self.m, self.plan = model, plan
        self.counters = dict(conv=0, bn=0, tconv=0)

    def _next(self, kind):
        n = self.counters[kind]
        self.counters[kind] = n + 1
        return f'{kind}{n}'

    # tensors are (TRef | None, channels) pairs so the variable pass needs no buffers
    def conv2d_bn(self, x, filters, k, activation='relu', dst=None):
        """multiresunet.py:31-60"""
        cname, bname = self._next('conv'), self._next('bn')
        ps, rng = self.m.params, self.m._ctx['rng']
        if self.plan is None:
            ps.add(f'{cname}/kernel', glorot_uniform(rng, (k, k, x[1], filters)))
            ps.add(f'{bname}/beta', np.zeros(filters, np.float32))
            ps.add(f'{bname}/moving_mean', np.zeros(filters, np.float32), trainable=False)
            ps.add(f'{bname}/moving_var', np.ones(filters, np.float32), trainable=False)
            return (None, filters)
        xr = x[0]
        y = dst or R.TRef(self.plan.new_buf(xr.h, xr.w, filters, cname))
        act = N.ACT_RELU if activation == 'relu' else N.ACT_NONE
        self.plan.add(_FoldedConv(self.plan, xr, y, cname, bname, k, act))
        return (y, filters)

    def full_bn(self, c):
        bname = self._next('bn')
        if self.plan is None:
            ps = self.m.params
            ps.add(f'{bname}/gamma', np.ones(c, np.float32))
            ps.add(f'{bname}/beta', np.zeros(c, np.float32))
            ps.add(f'{bname}/moving_mean', np.zeros(c, np.float32), trainable=False)
            ps.add(f'{bname}/moving_var', np.ones(c, np.float32), trainable=False)
            return None
        return _Affine(self.plan, bname, c)

    def tconv(self, x, filters, dst=None):
        name = self._next('tconv')
        if self.plan is None:
            ps, rng = self.m.params, self.m._ctx['rng']
            ps.add(f'{name}/kernel', glorot_uniform(rng, (2, 2, filters, x[1])))
            ps.add(f'{name}/bias', np.zeros(filters, np.float32))
            return (None, filters)
        self.plan.add(R.TConvOp(self.plan, x[0], dst, f'{name}/kernel', f'{name}/bias'))
        return (dst, filters)

    def mres_block(self, U, inp, alpha=1.67):
        """multiresunet.py:89-126"""
        W = alpha * U
        f1, f2, f3 = int(W * 0.167), int(W * 0.333), int(W * 0.5)
        ftot = f1 + f2 + f3
        shortcut = self.conv2d_bn(inp, ftot, 1, activation=None)
        if self.plan is None:
            c3 = self.conv2d_bn(inp, f1, 3)
            c5 = self.conv2d_bn(c3, f2, 3)
            self.conv2d_bn(c5, f3, 3)
            self.full_bn(ftot)
            self.full_bn(ftot)
            return (None, ftot)
        xr = inp[0]
        cat = self.plan.new_buf(xr.h, xr.w, ftot, 'mres_cat')          # concatenate([c3,c5,c7]) in place
        c3 = self.conv2d_bn(inp, f1, 3, dst=R.TRef(cat, 0, f1))
        c5 = self.conv2d_bn(c3, f2, 3, dst=R.TRef(cat, f1, f2))
        self.conv2d_bn(c5, f3, 3, dst=R.TRef(cat, f1 + f2, f3))
        bn1, bn2 = self.full_bn(ftot), self.full_bn(ftot)
        out = R.TRef(self.plan.new_buf(xr.h, xr.w, ftot, 'mres_out'))
        # out = BN2(relu(shortcut + BN1(cat)))
        self.plan.add(R.AddReluAffineOp(self.plan, shortcut[0], None, R.TRef(cat), bn1, bn2, out))
        return (out, ftot)

    def res_path(self, filters, length, inp, dst=None):
        """multiresunet.py:129-164"""
        out = inp
        for i in range(length):
            shortcut = self.conv2d_bn(out, filters, 1, activation=None)
            o = self.conv2d_bn(out, filters, 3)
            bn = self.full_bn(filters)
            if self.plan is None:
                out = (None, filters)
                continue
            xr = out[0]
            last = i == length - 1
            y = dst if (last and dst is not None) else R.TRef(self.plan.new_buf(xr.h, xr.w, filters, 'respath'))
            self.plan.add(R.AddReluAffineOp(self.plan, shortcut[0], None, o[0], None, bn, y))
            out = (y, filters)
        return out

    def pool(self, x):
        if self.plan is None:
            return x
        xr = x[0]
        y = R.TRef(self.plan.new_buf(xr.h // 2, xr.w // 2, x[1], 'pool'))
        self.plan.add(R.PoolOp(self.plan, xr, y))
        return (y, x[1])

    def graph(self, x):
        """multiresunet.py:180-221"""
        skips, cbufs = [], []
        for lvl, length in enumerate((4, 3, 2, 1)):
            U = 32 * 2 ** lvl
            b = self.mres_block(U, x)
            x = self.pool(b)
            dst = None
            if self.plan is not None:
                cb = self.plan.new_buf(b[0].h, b[0].w, 2 * U, f'up_concat{lvl}')   # [tconv (U) | respath (U)]
                cbufs.append(cb)
                dst = R.TRef(cb, U, U)
            skips.append(self.res_path(U, length, b, dst=dst))
        x = self.mres_block(32 * 16, x)
        for lvl in (3, 2, 1, 0):
            U = 32 * 2 ** lvl
            if self.plan is None:
                self.tconv(x, U)
                up = (None, 2 * U)
            else:
                self.tconv(x, U, dst=R.TRef(cbufs[lvl], 0, U))
                up = (R.TRef(cbufs[lvl]), 2 * U)
            x = self.mres_block(U, up)
        return x


def MultiResBlock(U, inp, alpha=1.67):
    """Exported by the reference registry (tf_models/__init__.py:2) as a Keras-functional helper; the
    B200 build lowers whole models, so the block is only reachable through ``MultiResUnet``."""
    raise NotImplementedError('MultiResBlock is emitted as part of MultiResUnet in this build')


class MultiResUnet(Model):
    """``MultiResUnet(height, width, n_channels)`` (multiresunet.py:167-223), inference path."""

    def __init__(self, height=None, width=None, n_channels=5, dtype=None, seed=0):
        super().__init__(dtype=dtype, seed=seed)
        self.configs = dict(height=height, width=width, n_channels=n_channels)
        self.n_channels = n_channels
        self.trainable_model = False
        self.input_shape = (None, height, width, n_channels)

    def get_config(self):
        return dict(self.configs)

    def _build_variables(self, input_shape):
        assert input_shape[-1] in (None, self.n_channels), 'n_channels is pinned by the config (multiresunet.yaml:5)'
        self.input_shape = (*input_shape[:3], self.n_channels)
        b = _Builder(self)
        x = b.graph((None, self.n_channels))
        b.conv2d_bn(x, 1, 1, activation='sigmoid')      # conv10, multiresunet.py:219 (head below)

    def _emit(self, plan):
        x = plan.input
        if plan.dtype != x.buf.dtype:
            xb = R.TRef(plan.new_buf(x.h, x.w, x.c, 'input_cast'))
            plan.add(R.ConvertOp(plan, x, xb))
            x = xb
        b = _Builder(self, plan)
        feats, c = b.graph((x, self.n_channels))
        # conv10 = conv2d_bn(.., 1, 1, 1, 'sigmoid'): 1x1 conv (no bias) -> BN(scale=False) -> sigmoid.
        # Lowered onto the fused head kernel with folded weights: logit = f.(K*s) + (beta - mean*s)
        cname, bname = b._next('conv'), b._next('bn')
        ps = self.params
        wf = torch.empty(c, dtype=torch.float32, device=plan.device)
        bf = torch.empty(1, dtype=torch.float32, device=plan.device)
        plan._head_fold = (wf, bf)

        def fold(train):
            s = torch.rsqrt(ps.view(f'{bname}/moving_var') + R.BN_EPSILON)
            torch.mul(ps.view(f'{cname}/kernel').view(-1), s, out=wf)
            torch.addcmul(ps.view(f'{bname}/beta'), ps.view(f'{bname}/moving_mean'), s, value=-1.0, out=bf)
        plan.add(R.CallbackOp(fold))
        plan.features = feats
        plan.head = None
        plan.head_forward = lambda: N.call('dnnca_head_fwd', N.stream_ptr(), feats.ct(), N.ptr(wf), N.ptr(bf),
                                           N.ptr(plan.logits), N.ptr(plan.probs))
