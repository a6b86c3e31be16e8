"""MultiResUnet with the reference's signature (``multiresunet.py:167-223``),
forward / inference only (BASELINE configs[4] is a forward sweep; SURVEY.md 2 "M8").

Inference-time lowering: ``conv2d_bn`` = Conv2D(no bias) -> BN(scale=False) -> act
(multiresunet.py:31-60) with moving statistics is a per-output-channel affine of the conv
result, i.e. conv with scaled kernels + a bias (exact: zero padding precedes the conv);
the block tails BN -> add -> relu -> BN (multiresunet.py:119-124, 148-150) run as one
``dnnca_add_relu_affine`` pass; the three chain convs of a MultiRes block write disjoint
channel ranges of one buffer (``concatenate``, multiresunet.py:119).

Tensor cores for the odd widths.  The block widths int(1.67*U*{.167,.333,.5}) are 8/17/26,
17/35/53, 35/71/106, 71/142/213, 142/284/427 -- not multiples of 8, so neither a 16-byte TMA row
nor a 16-byte epilogue store fits them.  Every logical tensor therefore lives in a buffer whose
concat segments are padded to multiples of 8 channels (51 -> 8|24|32 = 64, 105 -> 120, 212 -> 224,
426 -> 432, 853 -> 864); a ``Sym`` carries the logical->physical channel map.  The folded kernels
(``dnnca_fold_weights``) have zero rows / columns / biases at the holes, so holes hold exact zeros
through conv, pool, add-relu and ConvT and never touch a result.  All 56 convs and the 4 ConvT then
run on the tcgen05 implicit-GEMM kernels.

Weights are folded and packed ONCE per weight version: the first forward after ``set_weights`` /
``load_weights`` runs eagerly with the fp32 folded kernels (each fprop re-packs them to bf16 into its
layer's workspace); from then on every fprop passes ``w = NULL`` (prepacked, ``dnnca.h``) and the
captured graph holds the conv kernels only.
"""
from __future__ import annotations

import numpy as np
import torch

from ... import native as N
from ... import runtime as R
from ...keras_like import Model
from .components import glorot_uniform


def pad8(c):
    return (c + 7) // 8 * 8


class Sym:
    """A tensor of the reference graph while it is being walked: ``tref`` (None in the variable pass), its logical
    channel count and the segments ``[(physical offset inside the view, logical channels)]`` it occupies."""

    def __init__(self, builder, tref, c, segs=None):
        self.b, self.tref, self.c = builder, tref, c
        self.segs = segs if segs is not None else [(0, c)]

    @property
    def cphys(self):
        return self.tref.c if self.tref is not None else sum(pad8(n) for _, n in self.segs)

    def chmap(self):
        """int32 [cphys]: logical channel of every physical one, -1 for holes."""
        m = np.full(self.cphys, -1, np.int32)
        lo = 0
        for off, n in self.segs:
            m[off:off + n] = np.arange(lo, lo + n)
            lo += n
        assert lo == self.c
        return m


def _dev_map(plan, m):
    return torch.from_numpy(np.ascontiguousarray(m, np.int32)).to(plan.device)


class _FoldedConv(R.Op):
    """conv2d_bn at inference: y = act(conv(x, K*s) + t), s/t from the BN moving statistics, folded and laid out over
    the physical channels on the device (``dnnca_fold_weights``)."""

    def __init__(self, plan, x: Sym, y: R.TRef, out_map, conv, bn, ksize, act, cout):
        self.p, self.xs, self.x, self.y, self.conv, self.bn, self.k, self.act, self.cout = plan, x, x.tref, y, conv, bn, ksize, act, cout
        self.out_map_np = out_map
        self.wf = self.bf = self.ws = self.in_map = self.out_map = None

    def allocate(self, training):
        if self.wf is None:
            dev = self.p.device
            self.wf = torch.empty(self.k * self.k * self.x.c * self.y.c, dtype=torch.float32, device=dev)
            self.bf = torch.empty(self.y.c, dtype=torch.float32, device=dev)
            self.in_map = _dev_map(self.p, self.xs.chmap())
            self.out_map = _dev_map(self.p, self.out_map_np)
            self.ws = R.conv_workspace(self.p, self.k * self.k, [self.x], self.y, out_multiple=8)

    def fold(self):
        ps = self.p.params
        N.call('dnnca_fold_weights', N.stream_ptr(), ps.ptr(f'{self.conv}/kernel'), self.k * self.k, self.xs.c, self.cout, 0,
               N.ptr(self.in_map), self.x.c, N.ptr(self.out_map), self.y.c, None, ps.ptr(f'{self.bn}/beta'),
               ps.ptr(f'{self.bn}/moving_mean'), ps.ptr(f'{self.bn}/moving_var'), R.BN_EPSILON, N.ptr(self.wf), N.ptr(self.bf))

    def fwd(self, train):
        assert not train, 'MultiResUnet is forward/inference-only in this build'
        tc = self.ws is not None                 # tensor-core layer: packed once per weight version, then w = NULL
        if not (self.p.prepacked and tc):
            self.fold()
            if tc:
                N.call('dnnca_conv2d_prepack', N.stream_ptr(), self.x.ct(), None, N.ptr(self.wf), self.y.ct(), self.k,
                       *R.ws_args(self.ws))
        N.call('dnnca_conv2d_fprop', N.stream_ptr(), self.x.ct(), None, None if tc else N.ptr(self.wf), N.ptr(self.bf),
               self.y.ct(), self.k, self.act, 0.0, None, *R.ws_args(self.ws))


class _FoldedTConv(R.Op):
    """Conv2DTranspose 2x2/2 with bias, no BN (multiresunet.py:200-215) reading a physically padded input."""

    def __init__(self, plan, x: Sym, y: R.TRef, name, cout):
        self.p, self.xs, self.x, self.y, self.name, self.cout = plan, x, x.tref, y, name, cout
        self.wf = self.ws = self.in_map = None

    def allocate(self, training):
        if self.wf is None:
            self.wf = torch.empty(4 * self.y.c * self.x.c, dtype=torch.float32, device=self.p.device)
            self.in_map = _dev_map(self.p, self.xs.chmap())
            self.ws = R.conv_workspace(self.p, 4, [self.x], self.y)

    def fold(self):
        ps = self.p.params
        N.call('dnnca_fold_weights', N.stream_ptr(), ps.ptr(f'{self.name}/kernel'), 4, self.xs.c, self.cout, 1,
               N.ptr(self.in_map), self.x.c, None, self.y.c, None, None, None, None, 0.0, N.ptr(self.wf), None)

    def fwd(self, train):
        assert not train, 'MultiResUnet is forward/inference-only in this build'
        tc = self.ws is not None
        if not (self.p.prepacked and tc):
            self.fold()
            if tc:
                N.call('dnnca_convtranspose2x2_prepack', N.stream_ptr(), self.x.ct(), N.ptr(self.wf), self.y.ct(),
                       *R.ws_args(self.ws))
        N.call('dnnca_convtranspose2x2_fprop', N.stream_ptr(), self.x.ct(), None if tc else N.ptr(self.wf),
               self.p.params.ptr(f'{self.name}/bias'), self.y.ct(), None, *R.ws_args(self.ws))


class _Affine:
    """scale|shift [2*Cphys] of an inference BatchNormalization over the physical channels (holes: 0 | 0); refreshed
    together with the folded kernels (once per weight version)."""

    def __init__(self, plan, bn, sym: Sym, scale=True):
        self.p, self.bn, self.sym, self.scale = plan, bn, sym, scale
        self.buf = self.map = None

    def __call__(self):
        ps = self.p.params
        cp = self.sym.cphys
        if self.buf is None:
            self.buf = torch.empty(2 * cp, dtype=torch.float32, device=self.p.device)
            self.map = _dev_map(self.p, self.sym.chmap())
        if not self.p.prepacked:
            N.call('dnnca_bn_inference_params_mapped', N.stream_ptr(), cp, N.ptr(self.map),
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

    def conv2d_bn(self, x: Sym, filters, k, activation='relu', dst: R.TRef = None, out_segs=None):
        """multiresunet.py:31-60.  ``dst`` / ``out_segs``: place the output inside a wider (concat) buffer with the
        given logical->physical layout; default = a fresh buffer of pad8(filters) channels."""
        cname, bname = self._next('conv'), self._next('bn')
        ps, rng = self.m.params, self.m._ctx['rng']
        if self.plan is None:
            ps.add(f'{cname}/kernel', glorot_uniform(rng, (k, k, x.c, filters)))
            ps.add(f'{bname}/beta', np.zeros(filters, np.float32))
            ps.add(f'{bname}/moving_mean', np.zeros(filters, np.float32), trainable=False)
            ps.add(f'{bname}/moving_var', np.ones(filters, np.float32), trainable=False)
            return Sym(self, None, filters)
        xr = x.tref
        y = dst or R.TRef(self.plan.new_buf(xr.h, xr.w, pad8(filters), cname, zero=True))
        out = Sym(self, y, filters, out_segs)
        act = N.ACT_RELU if activation == 'relu' else N.ACT_NONE
        self.plan.add(_FoldedConv(self.plan, x, y, out.chmap(), cname, bname, k, act, filters))
        return out

    def full_bn(self, sym: Sym):
        bname = self._next('bn')
        if self.plan is None:
            ps, c = self.m.params, sym.c
            ps.add(f'{bname}/gamma', np.ones(c, np.float32))
            ps.add(f'{bname}/beta', np.zeros(c, np.float32))
            ps.add(f'{bname}/moving_mean', np.zeros(c, np.float32), trainable=False)
            ps.add(f'{bname}/moving_var', np.ones(c, np.float32), trainable=False)
            return None
        return _Affine(self.plan, bname, sym)

    def tconv(self, x: Sym, filters, dst: R.TRef = None):
        name = self._next('tconv')
        if self.plan is None:
            ps, rng = self.m.params, self.m._ctx['rng']
            ps.add(f'{name}/kernel', glorot_uniform(rng, (2, 2, filters, x.c)))
            ps.add(f'{name}/bias', np.zeros(filters, np.float32))
            return Sym(self, None, filters)
        assert dst.c == filters
        self.plan.add(_FoldedTConv(self.plan, x, dst, name, filters))
        return Sym(self, dst, filters)

    def mres_block(self, U, inp: Sym, alpha=1.67):
        """multiresunet.py:89-126"""
        W = alpha * U
        f1, f2, f3 = int(W * 0.167), int(W * 0.333), int(W * 0.5)
        ftot = f1 + f2 + f3
        if self.plan is None:
            self.conv2d_bn(inp, ftot, 1, activation=None)
            c3 = self.conv2d_bn(inp, f1, 3)
            c5 = self.conv2d_bn(c3, f2, 3)
            self.conv2d_bn(c5, f3, 3)
            cat = Sym(self, None, ftot)
            self.full_bn(cat)
            self.full_bn(cat)
            return Sym(self, None, ftot)
        xr = inp.tref
        p1, p2, p3 = pad8(f1), pad8(f2), pad8(f3)
        segs = [(0, f1), (p1, f2), (p1 + p2, f3)]                # logical 51 = 8|17|26 over physical 8|24|32
        cphys = p1 + p2 + p3
        # the shortcut shares the concat's physical layout (the add is elementwise over physical channels)
        sbuf = R.TRef(self.plan.new_buf(xr.h, xr.w, cphys, 'mres_shortcut', zero=True))
        shortcut = self.conv2d_bn(inp, ftot, 1, activation=None, dst=sbuf, out_segs=segs)
        cat = self.plan.new_buf(xr.h, xr.w, cphys, 'mres_cat', zero=True)          # concatenate([c3,c5,c7]) in place
        c3 = self.conv2d_bn(inp, f1, 3, dst=R.TRef(cat, 0, p1))
        c5 = self.conv2d_bn(c3, f2, 3, dst=R.TRef(cat, p1, p2))
        self.conv2d_bn(c5, f3, 3, dst=R.TRef(cat, p1 + p2, p3))
        cats = Sym(self, R.TRef(cat), ftot, segs)
        bn1, bn2 = self.full_bn(cats), self.full_bn(cats)
        out = R.TRef(self.plan.new_buf(xr.h, xr.w, cphys, 'mres_out', zero=True))
        # out = BN2(relu(shortcut + BN1(cat)))
        self.plan.add(R.AddReluAffineOp(self.plan, shortcut.tref, None, cats.tref, bn1, bn2, out))
        return Sym(self, out, ftot, segs)

    def res_path(self, filters, length, inp: Sym, dst: R.TRef = None):
        """multiresunet.py:129-164"""
        out = inp
        for i in range(length):
            shortcut = self.conv2d_bn(out, filters, 1, activation=None)
            o = self.conv2d_bn(out, filters, 3)
            if self.plan is None:
                self.full_bn(Sym(self, None, filters))
                out = Sym(self, None, filters)
                continue
            xr = out.tref
            last = i == length - 1
            y = dst if (last and dst is not None) else R.TRef(self.plan.new_buf(xr.h, xr.w, filters, 'respath'))
            ysym = Sym(self, y, filters)
            bn = self.full_bn(ysym)
            self.plan.add(R.AddReluAffineOp(self.plan, shortcut.tref, None, o.tref, None, bn, y))
            out = ysym
        return out

    def pool(self, x: Sym):
        if self.plan is None:
            return x
        xr = x.tref
        y = R.TRef(self.plan.new_buf(xr.h // 2, xr.w // 2, xr.c, 'pool'))
        self.plan.add(R.PoolOp(self.plan, xr, y))
        return Sym(self, y, x.c, x.segs)

    def graph(self, x: Sym):
        """multiresunet.py:180-221"""
        skips, cbufs = [], []
        for lvl, length in enumerate((4, 3, 2, 1)):
            U = 32 * 2 ** lvl
            b = self.mres_block(U, x)
            x = self.pool(b)
            dst = None
            if self.plan is not None:
                cb = self.plan.new_buf(b.tref.h, b.tref.w, 2 * U, f'up_concat{lvl}')   # [tconv (U) | respath (U)]
                cbufs.append(cb)
                dst = R.TRef(cb, U, U)
            skips.append(self.res_path(U, length, b, dst=dst))
        x = self.mres_block(32 * 16, x)
        for lvl in (3, 2, 1, 0):
            U = 32 * 2 ** lvl
            if self.plan is None:
                self.tconv(x, U)
                up = Sym(self, None, 2 * U)
            else:
                self.tconv(x, U, dst=R.TRef(cbufs[lvl], 0, U))
                up = Sym(self, R.TRef(cbufs[lvl]), 2 * U)
            x = self.mres_block(U, up)
        return x


def MultiResBlock(U, inp, alpha=1.67):
    """``MultiResBlock(U, inp, alpha)`` (multiresunet.py:89-126; exported by the reference registry,
    tf_models/__init__.py:2).  ``inp`` is a tensor of the graph being built (a ``Sym`` handed out by this module's
    builder, the counterpart of a Keras functional tensor); returns the block's output tensor."""
    if not isinstance(inp, Sym):
        raise TypeError('MultiResBlock expects a tensor of a MultiResUnet graph under construction (Sym)')
    return inp.b.mres_block(U, inp, alpha)


class MultiResUnet(Model):
    """``MultiResUnet(height, width, n_channels)`` (multiresunet.py:167-223).

    Two plans per input shape: inference (``model(x)``, ``predict``, ``evaluate``) runs the BatchNorm-folded, prepacked
    tensor-core plan above; ``train_step`` / ``fit`` / ``forward_backward`` / ``model(x, training=True)`` run the training
    plan of ``multires_train.py`` (batch statistics, full backward pass, the variables stay in the reference's shapes)."""

    def __init__(self, height=None, width=None, n_channels=5, dtype=None, seed=0):
        super().__init__(dtype=dtype, seed=seed)
        self.configs = dict(height=height, width=width, n_channels=n_channels)
        self.n_channels = n_channels
        self.input_shape = (None, height, width, n_channels)
        self._train_plans = 0           # > 0 while a training entry point is choosing its plan

    def get_config(self):
        return dict(self.configs)

    def _build_variables(self, input_shape):
        assert input_shape[-1] in (None, self.n_channels), 'n_channels is pinned by the config (multiresunet.yaml:5)'
        self.input_shape = (*input_shape[:3], self.n_channels)
        b = _Builder(self)
        x = b.graph(Sym(b, None, self.n_channels))
        b.conv2d_bn(x, 1, 1, activation='sigmoid')      # conv10, multiresunet.py:219 (head below)

    def _emit(self, plan):
        x = plan.input
        plan.prepacked = False
        plan.weights_version = None
        b = _Builder(self, plan)
        if plan.dtype != x.buf.dtype:
            # 16-byte pixels for the TMA: 5 modalities in an 8-channel buffer whose last three channels stay zero
            buf = plan.new_buf(x.h, x.w, pad8(x.c), 'input_cast', zero=True)
            plan.add(R.ConvertOp(plan, x, R.TRef(buf, 0, x.c)))
            xs = Sym(b, R.TRef(buf), self.n_channels)
        else:
            xs = Sym(b, x, self.n_channels)
        feats = b.graph(xs)
        # conv10 = conv2d_bn(.., 1, 1, 1, 'sigmoid'): 1x1 conv (no bias) -> BN(scale=False) -> sigmoid.
        # Lowered onto the fused head kernel with folded weights: logit = f.(K*s) + (beta - mean*s)
        cname, bname = b._next('conv'), b._next('bn')
        ps = self.params
        cp = feats.cphys
        wf = torch.empty(cp, dtype=torch.float32, device=plan.device)
        bf = torch.empty(1, dtype=torch.float32, device=plan.device)
        in_map = _dev_map(plan, feats.chmap())
        plan._head_fold = (wf, bf, in_map)

        def fold(train):
            if not plan.prepacked:
                N.call('dnnca_fold_weights', N.stream_ptr(), ps.ptr(f'{cname}/kernel'), 1, feats.c, 1, 0, N.ptr(in_map), cp,
                       None, 1, None, ps.ptr(f'{bname}/beta'), ps.ptr(f'{bname}/moving_mean'), ps.ptr(f'{bname}/moving_var'),
                       R.BN_EPSILON, N.ptr(wf), N.ptr(bf))
        plan.add(R.CallbackOp(fold))
        plan.features = feats.tref
        plan.head = None
        plan.head_forward = lambda: N.call('dnnca_head_fwd', N.stream_ptr(), feats.tref.ct(), N.ptr(wf), N.ptr(bf),
                                           N.ptr(plan.logits), N.ptr(plan.probs))

    # ---- plan selection --------------------------------------------------------------------------------------
    def _plan(self, batch, height, width, want_input_grad=False):
        if want_input_grad:
            # callbacks.py:290-299: the training plan's op list run with the moving statistics + its dgrad chain
            key = (batch, height, width, 'dx')
            if key not in self._plans:
                from .multires_train import emit_training_plan
                super()._plan(batch, height, width)
                self._plans[key] = emit_training_plan(self, batch, height, width, want_input_grad=True)
            return self._plans[key]
        if not self._train_plans:
            return super()._plan(batch, height, width)
        key = (batch, height, width, 'train')
        if key not in self._plans:
            from .multires_train import emit_training_plan
            super()._plan(batch, height, width)          # builds / materialises the variables (and the inference plan)
            self._plans[key] = emit_training_plan(self, batch, height, width)
        return self._plans[key]

    def training_plan(self, batch, height, width):
        with self._training():
            return self._plan(batch, height, width)

    def _training(self):
        model = self

        class _Ctx:
            def __enter__(self):
                model._train_plans += 1

            def __exit__(self, *exc):
                model._train_plans -= 1
        return _Ctx()

    def _before_inference(self, plan):
        """Folded / packed weights are refreshed (one eager pass over the batch already staged in ``plan.x_in``) whenever
        the variables changed since the last call; the steady state replays a graph of prepacked tensor-core convs."""
        if self._train_plans:
            return
        if plan.weights_version != self.params.version:
            plan.prepacked = False
            plan.forward(False)                    # folds every kernel / affine and leaves the bf16 packings behind
            plan.weights_version = self.params.version
        plan.prepacked = plan.dtype == torch.bfloat16

    def __call__(self, x, training=False):
        if training:
            with self._training():
                self.params.version += 1        # the call moves the BatchNorm averages: inference plans re-fold
                return super().__call__(x, training=True)
        return super().__call__(x, training=False)

    def train_step(self, x, y, lr=None):
        with self._training():
            return super().train_step(x, y, lr)

    def input_gradient(self, x):
        with self._training():                  # (keeps _before_inference away from the padded plan)
            return super().input_gradient(x)

    def forward_backward(self, x, y):
        with self._training():
            self.params.version += 1            # (moving statistics change)
            return super().forward_backward(x, y)

    def prefetch(self, x, y):
        with self._training():
            return super().prefetch(x, y)
