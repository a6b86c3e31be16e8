"""Training plan of ``MultiResUnet`` (``multiresunet.py:31-223`` under ``engine.py:286`` ``model.fit``).

The reference trains ``configs/multiresunet.yaml`` like every other config: every ``conv2d_bn`` is
Conv2D(no bias) -> BatchNormalization(scale=False) with BATCH statistics -> activation, the block tails are
BatchNormalization -> add -> relu -> BatchNormalization (multiresunet.py:119-124), and the gradient flows through all of it.

Channel padding.  The block widths (8/17/26, 17/35/53, ...) are not multiples of 16, which the tcgen05 conv kernels
(fprop, dgrad AND wgrad) want.  The training plan therefore computes on *physical* tensors whose concat segments are padded
to multiples of 16 channels, and on physical copies of every variable (``ParamStore`` ``phys``) with zeros at the holes:

    step:  phys variables <- gather(logical variables)          one ``dnnca_gather_f32`` (holes: 0)
           forward / backward on the physical plan               the conv / BN / pool ops the U-Nets use
           logical gradients <- gather(phys gradients)           one ``dnnca_gather_f32``
           logical moving statistics <- gather(phys ones)        one ``dnnca_gather_f32``
           fused Adam on the logical variables                   unchanged (checkpoints, get_weights: reference shapes)

Holes stay exact zeros through the whole pass: a hole output column has zero weights (conv output 0), its BatchNorm has
mean 0 / gamma 0 / beta 0 (output 0), relu / add / pool / ConvT keep 0; hole input rows have zero weights, so hole
gradients never reach a real channel, and the gradients of hole weights are never gathered.

Fan-out.  Tensors with several consumers (block input -> 1x1 shortcut and 3x3 chain; conv3x3 -> conv5x5 and the concat;
block output -> pool and ResPath) get ONE direct writer of their gradient -- the consumer that runs first in the backward
pass -- the other consumers' dgrads go to a scratch tensor that ``dnnca_accumulate`` adds (the max-pool adds in place,
``dnnca_maxpool2x2_bwd(dskip=)``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ... import native as N
from ... import runtime as R

PAD = 16
RELU = (N.ACT_RELU, 0.0)


def padc(c):
    return (c + PAD - 1) // PAD * PAD


class TSym:
    """A tensor of the graph: its TRef over a physical buffer, the logical channel count and the segments
    ``[(physical offset inside the view, logical channels)]``."""

    def __init__(self, tref, c, segs=None):
        self.tref, self.c = tref, c
        self.segs = segs if segs is not None else [(0, c)]

    @property
    def cphys(self):
        return self.tref.c

    def pos(self):
        """int [c]: physical channel (inside the view) of every logical channel"""
        return np.concatenate([np.arange(off, off + n) for off, n in self.segs]).astype(np.int64)


class _Alias(R.TRef):
    """Same storage (and activation mask) as ``base``, but its gradient view comes from ``gfn`` -- another tensor's
    gradient (shortcut + BN1 output both receive the gradient of their sum) or a scratch tensor (fan-out)."""

    def __init__(self, base: R.TRef, gfn, own_grad=True):
        super().__init__(base.buf, base.coff, base.c)
        self.act = base.act
        self.needs_grad = base.needs_grad and own_grad      # own_grad=False: no gradient buffer behind THIS storage
        self.skip_consumed = base.skip_consumed
        self._gfn = gfn

    def gct(self):
        return self._gfn()


class BNActOp(R.BNOp):
    """BatchNormalization(scale=False) -> Activation (conv2d_bn, multiresunet.py:53-58) in training: the apply pass carries
    the activation; the gradient arriving in ``y`` is already masked by its writers (``TRef.mask_args``), so the backward
    pass is the plain BatchNorm one."""

    def __init__(self, plan, x, y, prefix, stats, act):
        super().__init__(plan, x, y, prefix, stats, scale=False, fused_stats=True)
        self.folded = True                 # the base class skips its own (activation-free) apply pass
        self.post_act = act or (N.ACT_NONE, 0.0)
        if act:
            y.act = act

    def fwd(self, train):
        super().fwd(train)
        N.call('dnnca_bn_apply_act', N.stream_ptr(), self.x.ct(), N.ptr(self.ss), self.y.ct(), self.post_act[0], self.post_act[1])


class AddReluOp(R.Op):
    """t = relu(a + affine_b(b)) (multiresunet.py:121-122, 148-149); ``bn_b``: the BatchNorm whose scale|shift applies
    to b (block tail) or None (ResPath).  Backward: a's and the BatchNorm output's gradient ARE t's (aliases); for the
    ResPath, b = relu(BN(conv)) needs its own mask: db = dt * [b > 0]."""

    def __init__(self, plan, a, b, bn_b, t, mask_b=False):
        self.p, self.a, self.b, self.bn_b, self.t, self.mask_b = plan, a, b, bn_b, t, mask_b
        t.act = RELU

    def fwd(self, train):
        N.call('dnnca_add_relu_affine', N.stream_ptr(), self.a.ct(), None, self.b.ct(),
               N.ptr(self.bn_b.ss) if self.bn_b is not None else None, None, self.t.ct())

    def bwd(self):
        if self.mask_b:
            N.call('dnnca_act_bwd', N.stream_ptr(), self.b.ct(), self.t.gct(), self.b.gct(), N.ACT_RELU, 0.0)

    bwd_input = bwd


class AccumOp(R.Op):
    """Placed in FRONT of a consumer whose dgrad went to ``scratch``: runs right after it in the backward pass."""

    def __init__(self, plan, scratch: R.TRef, dst: R.TRef):
        self.p, self.scratch, self.dst = plan, scratch, dst

    def fwd(self, train):
        pass

    def bwd(self):
        N.call('dnnca_accumulate', N.stream_ptr(), self.scratch.ct(), self.dst.gct())

    bwd_input = bwd


class MultiResTrainPlan(R.Plan):
    """Physical (channel-padded) plan + the gathers between the logical variables and their padded copies."""

    def __init__(self, logical: R.ParamStore, batch, height, width, channels, dtype, device, want_input_grad=False):
        self.logical = logical
        self.phys = R.ParamStore()
        super().__init__(self.phys, batch, height, width, channels, dtype, device, want_input_grad=want_input_grad)
        self.links = {}                    # variable name -> per-axis physical positions (np.ix_ arguments) or None
        self._scratch = {}
        self.maps = None

    # ---- variables ---------------------------------------------------------------------------------
    def add_var(self, name, phys_shape, index=None):
        spec = self.logical.specs[name]
        self.phys.add(name, np.zeros(phys_shape, np.float32), trainable=spec['trainable'])
        self.links[name] = index

    def scratch(self, like: R.TRef):
        key = (like.h, like.w, like.c)
        if key not in self._scratch:
            self._scratch[key] = R.TRef(R.Buf(self, like.n, like.h, like.w, like.c, f'fanout_scratch{key}'))
            self._scratch[key].needs_grad = False
        return self._scratch[key]

    def build_maps(self):
        """int32 index maps between the flat logical and physical buffers (parameters / gradients and BatchNorm state)."""
        lg, ph = self.logical, self.phys
        ph.materialize(self.device)
        n = dict(lt=max(lg.n_trainable, 4), ls=max(lg.n_state, 4), pt=max(ph.n_trainable, 4), ps=max(ph.n_state, 4))
        p2l_t, p2l_s = np.full(n['pt'], -1, np.int32), np.full(n['ps'], -1, np.int32)
        l2p_t, l2p_s = np.full(n['lt'], -1, np.int32), np.full(n['ls'], -1, np.int32)
        for name, index in self.links.items():
            ls_, ps_ = lg.specs[name], ph.specs[name]
            lidx = (np.arange(ls_['numel'], dtype=np.int64) + ls_['offset']).reshape(ls_['shape'])
            pidx = (np.arange(ps_['numel'], dtype=np.int64) + ps_['offset']).reshape(ps_['shape'])
            sel = pidx[np.ix_(*index)] if index is not None else pidx
            assert sel.shape == lidx.shape, (name, sel.shape, lidx.shape)
            p2l, l2p = (p2l_t, l2p_t) if ls_['trainable'] else (p2l_s, l2p_s)
            p2l[sel.ravel()] = lidx.ravel()
            l2p[lidx.ravel()] = sel.ravel()
        dev = self.device
        self.maps = {k: torch.from_numpy(v).to(dev) for k, v in dict(p2l_t=p2l_t, p2l_s=p2l_s, l2p_t=l2p_t, l2p_s=l2p_s).items()}

    def _gather(self, src, idx, dst):
        N.call('dnnca_gather_f32', N.stream_ptr(), N.ptr(src), N.ptr(idx), idx.numel(), N.ptr(dst))

    # ---- step protocol (keras_like.Model launch sequences) -------------------------------------------
    def refresh_variables(self):
        lg, ph, m = self.logical, self.phys, self.maps
        self._gather(lg.params, m['p2l_t'], ph.params)
        self._gather(lg.state, m['p2l_s'], ph.state)

    def zero_step_state(self):
        """Start of a training-mode launch sequence: gradients / statistics zeroed, physical variables refreshed."""
        super().zero_step_state()
        self.logical.grads_full.zero_()
        self.refresh_variables()

    def forward(self, train=False):
        if not train:          # the input-gradient chain (inference-mode forward) has no zero_step_state ahead of it
            self.refresh_variables()
        super().forward(train)

    def end_backward(self):
        lg, ph, m = self.logical, self.phys, self.maps
        self._gather(ph.grads, m['l2p_t'], lg.grads)
        self._gather(ph.state, m['l2p_s'], lg.state)

    def ready_frontier(self):
        """Data parallelism: the logical gradients exist only after ``end_backward`` -- no bucket leaves earlier."""
        n = self.logical.grads_full.numel()
        self.pending_before_backward = n
        return [n] * len(self.ops)

    # ---- head: conv10 = Conv2D(1, 1x1, no bias) -> BN(scale=False) -> sigmoid (multiresunet.py:219) ------------------
    def set_head(self, feats: TSym, cname, bname):
        self.features = feats.tref
        self.head = None
        self.head_names = (cname, bname)
        self.head_stats = R.BNStats(self, 1)

    def _head_buffers(self):
        if getattr(self, 'hz', None) is None:
            B, H, W, dev = self.batch, self.height, self.width, self.device
            mk = lambda: torch.zeros(B, H, W, 1, dtype=torch.float32, device=dev)
            self.hz, self.hzg, self.hl, self.hlg = mk(), mk(), mk(), mk()      # conv output, its gradient, BN output, its gradient
            self.hss = torch.zeros(2, dtype=torch.float32, device=dev)
            self.hmi = torch.zeros(2, dtype=torch.float32, device=dev)
            self.hone = torch.ones(4, dtype=torch.float32, device=dev)
            self.hzero = torch.zeros(4, dtype=torch.float32, device=dev)
            self.hdump = torch.zeros(8, dtype=torch.float32, device=dev)
            self._hv = {k: N.tensor_view(getattr(self, k)) for k in ('hz', 'hzg', 'hl', 'hlg')}

    def _head_logits(self, train):
        """z = f . k (fp32), BatchNorm over its single channel with the batch statistics -> the logits tensor ``hl``"""
        self._head_buffers()
        ps, s = self.phys, N.stream_ptr()
        cname, bname = self.head_names
        v = {k: C.byref(t) for k, t in self._hv.items()}
        N.call('dnnca_head_fwd', s, self.features.ct(), ps.ptr(f'{cname}/kernel'), None, N.ptr(self.hz), None)
        if train:
            N.call('dnnca_channel_stats', s, v['hz'], self.head_stats.fwd_ptr())
            N.call('dnnca_bn_finalize', s, self.head_stats.fwd_ptr(), self.batch * self.height * self.width, 1, None,
                   ps.ptr(f'{bname}/beta'), R.BN_MOMENTUM, R.BN_EPSILON, ps.ptr(f'{bname}/moving_mean'),
                   ps.ptr(f'{bname}/moving_var'), N.ptr(self.hss), N.ptr(self.hmi))
        else:
            N.call('dnnca_bn_inference_params', s, 1, None, ps.ptr(f'{bname}/beta'), R.BN_EPSILON,
                   ps.ptr(f'{bname}/moving_mean'), ps.ptr(f'{bname}/moving_var'), N.ptr(self.hss))
        N.call('dnnca_bn_apply', s, v['hz'], N.ptr(self.hss), v['hl'])
        return v

    def _head_folded(self):
        """conv10 + its BatchNorm with the MOVING statistics as one weight vector + bias (``dnnca_fold_weights``)"""
        ps = self.phys
        cname, bname = self.head_names
        cp = self.features.c
        if getattr(self, 'hwf', None) is None:
            self.hwf = torch.empty(cp, dtype=torch.float32, device=self.device)
            self.hbf = torch.empty(1, dtype=torch.float32, device=self.device)
        N.call('dnnca_fold_weights', N.stream_ptr(), ps.ptr(f'{cname}/kernel'), 1, cp, 1, 0, None, cp, None, 1, None,
               ps.ptr(f'{bname}/beta'), ps.ptr(f'{bname}/moving_mean'), ps.ptr(f'{bname}/moving_var'), R.BN_EPSILON,
               N.ptr(self.hwf), N.ptr(self.hbf))

    def head_forward(self):
        """``model(x, training=True)``: batch statistics, moving averages updated, no gradients.  On the input-gradient
        plan (callbacks.py:290-299: inference-mode forward): moving statistics, folded into the head weights."""
        if self.want_input_grad:
            self._head_folded()
            N.call('dnnca_head_fwd', N.stream_ptr(), self.features.ct(), N.ptr(self.hwf), N.ptr(self.hbf), N.ptr(self.logits),
                   N.ptr(self.probs))
            return
        v = self._head_logits(self.train_bn)
        N.call('dnnca_head_fwd', N.stream_ptr(), v['hl'], N.ptr(self.hone), None, N.ptr(self.logits), N.ptr(self.probs))
        lg, ph, m = self.logical, self.phys, self.maps
        self._gather(ph.state, m['l2p_s'], lg.state)

    def head_input_grad(self):
        f = self.features
        act = f.act or (N.ACT_NONE, 0.0)
        N.call('dnnca_head_input_grad', N.stream_ptr(), f.ct(), N.ptr(self.hwf), N.ptr(self.hbf), f.gct(), act[0], act[1])

    def head_loss(self, loss_cfg: N.LossConfig, with_grads=True):
        ps, s = self.phys, N.stream_ptr()
        cname, bname = self.head_names
        self.per_sample.zero_()
        if not loss_cfg.has_weight:
            N.call('dnnca_label_stats_init', s, N.ptr(self.lstats))
            N.call('dnnca_label_stats', s, N.ptr(self.y_in), self.y_in.numel(), N.ptr(self.lstats))
        v = self._head_logits(True)
        # weighted BCE on the logits tensor: the fused head kernel with one feature of weight 1 (its df = dlogit)
        N.call('dnnca_head_bce_fwd_bwd', s, v['hl'], N.ptr(self.hone), None, N.ptr(self.y_in), N.ptr(self.lstats),
               C.byref(loss_cfg), N.ptr(self.logits), N.ptr(self.probs), N.ptr(self.per_sample),
               v['hlg'] if with_grads else None, N.ACT_NONE, 0.0, N.ptr(self.hdump), N.ptr(self.hdump[4:]))
        if not with_grads:
            return
        N.call('dnnca_bn_bwd_reduce', s, v['hz'], v['hlg'], N.ptr(self.hmi), self.head_stats.bwd_ptr())
        N.call('dnnca_bn_bwd_apply', s, v['hz'], v['hlg'], N.ptr(self.hmi), None, self.head_stats.bwd_ptr(), v['hzg'],
               N.ACT_NONE, 0.0, None, ps.gptr(f'{bname}/beta'))
        f = self.features
        act = f.act or (N.ACT_NONE, 0.0)
        N.call('dnnca_head_conv_bwd', s, f.ct(), ps.ptr(f'{cname}/kernel'), N.ptr(self.hzg), f.gct(), act[0], act[1],
               ps.gptr(f'{cname}/kernel'))


class TrainBuilder:
    """Walks the reference graph in the SAME order as the variable pass of ``multiresunet._Builder`` (the names
    conv<i> / bn<i> / tconv<i> follow creation order) and emits the training ops on physical tensors."""

    def __init__(self, plan: MultiResTrainPlan):
        self.plan = plan
        self.counters = dict(conv=0, bn=0, tconv=0)

    def _next(self, kind):
        n = self.counters[kind]
        self.counters[kind] = n + 1
        return f'{kind}{n}'

    # ---- layers ----------------------------------------------------------------------------------------
    def conv2d_bn(self, x: TSym, xref: R.TRef, filters, k, activation='relu', dst: R.TRef = None):
        """Conv2D(use_bias=False) -> BN(scale=False, batch statistics) -> activation (multiresunet.py:31-60).
        ``xref``: the TRef the conv reads through (x.tref or a fan-out alias of it); ``dst``: a slice of a concat buffer."""
        plan = self.plan
        cname, bname = self._next('conv'), self._next('bn')
        cp = padc(filters)
        xr = x.tref
        z = R.TRef(plan.new_buf(xr.h, xr.w, cp, f'{cname}_z'))
        y = dst if dst is not None else R.TRef(plan.new_buf(xr.h, xr.w, cp, f'{cname}_y'))
        assert y.c == cp
        out = np.arange(filters)
        plan.add_var(f'{cname}/kernel', (k, k, xr.c, cp), (np.arange(k), np.arange(k), x.pos(), out))
        for v in ('beta', 'moving_mean', 'moving_var'):
            plan.add_var(f'{bname}/{v}', (cp,), (out,))
        st = R.BNStats(plan, cp)
        plan.add(R.ConvOp(plan, xref, z, f'{cname}/kernel', None, k, None, stats=st))
        plan.add(BNActOp(plan, z, y, bname, st, RELU if activation == 'relu' else None))
        return TSym(y, filters)

    def full_bn_vars(self, sym: TSym):
        bname = self._next('bn')
        for v in ('gamma', 'beta', 'moving_mean', 'moving_var'):
            self.plan.add_var(f'{bname}/{v}', (sym.cphys,), (sym.pos(),))
        return bname

    def fan_in(self, x: TSym, direct: bool):
        """The TRef a consumer of ``x`` reads through: x's own (this consumer's dgrad writes the gradient) or a scratch
        alias + the AccumOp that adds it (must be emitted BEFORE the consumer)."""
        if direct or not x.tref.needs_grad:
            return x.tref
        sc = self.plan.scratch(x.tref)
        self.plan.add(AccumOp(self.plan, sc, x.tref))
        return _Alias(x.tref, sc.ct)

    def mres_block(self, U, inp: TSym, alpha=1.67):
        """multiresunet.py:89-126"""
        plan = self.plan
        W = alpha * U
        f1, f2, f3 = int(W * 0.167), int(W * 0.333), int(W * 0.5)
        ftot = f1 + f2 + f3
        p1, p2, p3 = padc(f1), padc(f2), padc(f3)
        segs = [(0, f1), (p1, f2), (p1 + p2, f3)]
        cphys = p1 + p2 + p3
        xr = inp.tref
        t = R.TRef(plan.new_buf(xr.h, xr.w, cphys, 'mres_t'))            # relu(shortcut + BN1(cat))
        # shortcut (1x1, no activation) laid out like the concat: its kernel columns sit at the concat's positions
        cname, bname = self._next('conv'), self._next('bn')
        cat_pos = np.concatenate([np.arange(o, o + n) for o, n in segs])
        zs = R.TRef(plan.new_buf(xr.h, xr.w, cphys, f'{cname}_z'))
        sbuf = R.TRef(plan.new_buf(xr.h, xr.w, cphys, f'{cname}_y'))
        plan.add_var(f'{cname}/kernel', (1, 1, xr.c, cphys), (np.arange(1), np.arange(1), inp.pos(), cat_pos))
        for v in ('beta', 'moving_mean', 'moving_var'):
            plan.add_var(f'{bname}/{v}', (cphys,), (cat_pos,))
        st = R.BNStats(plan, cphys)
        plan.add(R.ConvOp(plan, self.fan_in(inp, direct=False), zs, f'{cname}/kernel', None, 1, None, stats=st))
        plan.add(BNActOp(plan, zs, _Alias(sbuf, t.gct, own_grad=False), bname, st, None))     # d(shortcut) = d(t)
        # 3x3 chain writing the concat buffer in place
        cat = plan.new_buf(xr.h, xr.w, cphys, 'mres_cat')
        c3 = self.conv2d_bn(inp, inp.tref, f1, 3, dst=R.TRef(cat, 0, p1))            # direct writer of d(inp)
        c5 = self.conv2d_bn(c3, self.fan_in(c3, direct=False), f2, 3, dst=R.TRef(cat, p1, p2))
        self.conv2d_bn(c5, self.fan_in(c5, direct=False), f3, 3, dst=R.TRef(cat, p1 + p2, p3))
        catr = R.TRef(cat)
        catr.act = RELU
        cats = TSym(catr, ftot, segs)
        # BN1 -> add -> relu -> BN2 (multiresunet.py:120-124)
        b1, b2 = self.full_bn_vars(cats), self.full_bn_vars(cats)
        bn1 = R.BNOp(plan, catr, _Alias(catr, t.gct), b1, R.BNStats(plan, cphys), scale=True, fused_stats=False)
        bn1.folded = True                       # its apply pass is the affine of the add-relu kernel
        plan.add(bn1)
        plan.add(AddReluOp(plan, sbuf, catr, bn1, t))
        out = R.TRef(plan.new_buf(xr.h, xr.w, cphys, 'mres_out'))
        plan.add(R.BNOp(plan, t, out, b2, R.BNStats(plan, cphys), scale=True, fused_stats=False))
        return TSym(out, ftot, segs)

    def res_path(self, filters, length, inp: TSym, dst: R.TRef):
        """multiresunet.py:129-164; the last stage writes ``dst`` (the skip half of the decoder's concat buffer)."""
        plan = self.plan
        out = inp
        for i in range(length):
            xr = out.tref
            t = R.TRef(plan.new_buf(xr.h, xr.w, filters, 'respath_t'))
            # shortcut: second writer of d(out) -> scratch + accumulate
            cname, bname = self._next('conv'), self._next('bn')
            zs = R.TRef(plan.new_buf(xr.h, xr.w, filters, f'{cname}_z'))
            sbuf = R.TRef(plan.new_buf(xr.h, xr.w, filters, f'{cname}_y'))
            plan.add_var(f'{cname}/kernel', (1, 1, xr.c, filters), (np.arange(1), np.arange(1), out.pos(), np.arange(filters)))
            for v in ('beta', 'moving_mean', 'moving_var'):
                plan.add_var(f'{bname}/{v}', (filters,), None)
            st = R.BNStats(plan, filters)
            plan.add(R.ConvOp(plan, self.fan_in(out, direct=False), zs, f'{cname}/kernel', None, 1, None, stats=st))
            plan.add(BNActOp(plan, zs, _Alias(sbuf, t.gct, own_grad=False), bname, st, None))
            o = self.conv2d_bn(out, out.tref, filters, 3)
            plan.add(AddReluOp(plan, sbuf, o.tref, None, t, mask_b=True))
            bn = self.full_bn_vars(TSym(t, filters))
            y = dst if i == length - 1 else R.TRef(plan.new_buf(xr.h, xr.w, filters, 'respath'))
            plan.add(R.BNOp(plan, t, y, bn, R.BNStats(plan, filters), scale=True, fused_stats=False))
            out = TSym(y, filters)
        return out

    def pool(self, x: TSym):
        plan = self.plan
        xr = x.tref
        xr.skip_consumed = True           # the ResPath's dgrads write d(x) first, the pool adds in place
        y = R.TRef(plan.new_buf(xr.h // 2, xr.w // 2, xr.c, 'pool'))
        plan.add(R.PoolOp(plan, xr, y))
        return TSym(y, x.c, x.segs)

    def tconv(self, x: TSym, filters, dst: R.TRef):
        plan = self.plan
        name = self._next('tconv')
        assert dst.c == filters and filters % PAD == 0
        plan.add_var(f'{name}/kernel', (2, 2, filters, x.cphys), (np.arange(2), np.arange(2), np.arange(filters), x.pos()))
        plan.add_var(f'{name}/bias', (filters,), None)
        plan.add(R.TConvOp(plan, x.tref, dst, f'{name}/kernel', f'{name}/bias'))

    def graph(self, x: TSym):
        """multiresunet.py:180-221"""
        plan = self.plan
        cbufs = []
        for lvl, length in enumerate((4, 3, 2, 1)):
            U = 32 * 2 ** lvl
            b = self.mres_block(U, x)
            x = self.pool(b)
            cb = plan.new_buf(b.tref.h, b.tref.w, 2 * U, f'up_concat{lvl}')          # [tconv (U) | respath (U)]
            cbufs.append(cb)
            self.res_path(U, length, b, dst=R.TRef(cb, U, U))
        x = self.mres_block(32 * 16, x)
        for lvl in (3, 2, 1, 0):
            U = 32 * 2 ** lvl
            self.tconv(x, U, dst=R.TRef(cbufs[lvl], 0, U))
            x = self.mres_block(U, TSym(R.TRef(cbufs[lvl]), 2 * U))
        return x


def emit_training_plan(model, batch, height, width, device=None, want_input_grad=False):
    """The training plan of ``model`` (a ``MultiResUnet`` whose logical variables are materialised).  ``device``: only the
    CPU-side structure test passes one (the plan is then inspected, never launched)."""
    nch = model.n_channels
    plan = MultiResTrainPlan(model.params, batch, height, width, nch, model.compute_dtype, device or model.device,
                             want_input_grad=want_input_grad)
    b = TrainBuilder(plan)
    x = plan.input
    buf = plan.new_buf(x.h, x.w, padc(nch), 'input_cast', zero=True)        # modalities in a 16-channel pixel, rest zero
    plan.add(R.ConvertOp(plan, x, R.TRef(buf, 0, nch)))
    xin = R.TRef(buf)
    xin.needs_grad = plan.want_input_grad          # callbacks.py:290-299 asks for d(output)/d(input); training never does
    feats = b.graph(TSym(xin, nch))
    cname, bname = b._next('conv'), b._next('bn')
    plan.add_var(f'{cname}/kernel', (1, 1, feats.cphys, 1), (np.arange(1), np.arange(1), feats.pos(), np.arange(1)))
    for v in ('beta', 'moving_mean', 'moving_var'):
        plan.add_var(f'{bname}/{v}', (1,), None)
    plan.set_head(feats, cname, bname)
    assert set(plan.links) == set(model.params.specs), 'training plan and variable pass disagree on the variables'
    plan.build_maps()
    return plan
