"""Static execution plan for the conv-stack hot path.

The reference runs its Keras layers through TensorFlow's tracing compiler
(``@tf.function`` on every ``call``, components.py:77,158,235,314) and
``tf.GradientTape``.  Here a model is lowered ONCE per (batch, H, W) into a flat
list of ops over pre-allocated NHWC buffers in HBM:

* every activation lives in a ``Buf``; a ``TRef`` is a channel-slice view of one
  (``dnnca_tensor_t``), so ``tf.concat`` (components.py:164, unet.py:187) is just
  two producers writing disjoint channel ranges of the same buffer;
* gradients mirror the activation buffers; each op knows the C-ABI calls of its
  forward and of its backward (forward list reversed = the tape);
* all parameters sit in one flat fp32 buffer (+ one flat gradient buffer, which
  is what the data-parallel all-reduce and the fused Adam see);
* the whole training step (zeroing, forward, loss, backward, Adam) is a fixed
  launch sequence on one stream -> captured in a CUDA graph and replayed.

Only libdnnca kernels do arithmetic; torch supplies device memory, streams and
graph capture.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np
import torch

from . import native as N

BN_MOMENTUM = 0.99   # keras BatchNormalization defaults (components.py:57 passes none)
BN_EPSILON = 1e-3


# ----------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------
class ParamStore:
    """Named variables packed into flat fp32 device buffers.

    trainable -> ``params`` (+ ``grads``, Adam ``m``/``v``); non-trainable (BN moving
    statistics) -> ``state``.  Names follow the oracle's convention so weights can be
    exchanged with ``get_weights`` / ``set_weights``.
    """

    def __init__(self):
        self.specs: 'OrderedDict[str, dict]' = OrderedDict()
        self.device = None
        self.params = self.grads = self.m = self.v = self.state = self.l2 = None
        self._views = {}
        self._gviews = {}
        self.n_trainable = 0
        self.n_state = 0
        self.version = 0          # bumped whenever variables are (re)written: inference plans cache folded / packed weights

    def add(self, name, array, trainable=True, l2=0.0):
        assert self.device is None, 'variables must be created before the store is materialised'
        assert name not in self.specs, f'duplicate variable {name}'
        array = np.ascontiguousarray(array, dtype=np.float32)
        self.specs[name] = dict(shape=array.shape, trainable=trainable, init=array, l2=float(l2 or 0.0))

    def __contains__(self, name):
        return name in self.specs

    def materialize(self, device):
        if self.device is not None:
            return
        self.device = device
        off_t = off_s = 0
        for name, s in self.specs.items():
            n = int(np.prod(s['shape'])) if len(s['shape']) else 1
            # 4-float (16 B) alignment of every variable inside the flat buffers
            if s['trainable']:
                s['offset'] = off_t
                off_t += (n + 3) // 4 * 4
            else:
                s['offset'] = off_s
                off_s += (n + 3) // 4 * 4
            s['numel'] = n
        self.n_trainable, self.n_state = off_t, off_s
        self.params = torch.zeros(max(off_t, 4), dtype=torch.float32, device=device)
        # gradient buffer = [n_trainable gradients | 4-float tail]; tail[0] is the step's loss scalar
        # (dnnca_loss_total), so that it is zeroed with the gradients and, under data parallelism, travels in the
        # same SUM all-reduce (keras reports the global mean loss under MirroredStrategy)
        self.grads_full = torch.zeros(max(off_t, 4) + 4, dtype=torch.float32, device=device)
        self.grads = self.grads_full[:max(off_t, 4)]
        self.loss_slot = self.grads_full[max(off_t, 4):max(off_t, 4) + 1]       # where the step's loss is READ
        self.loss_in = self.loss_slot                                             # where dnnca_loss_total WRITES it
        self.m = torch.zeros_like(self.params)
        self.v = torch.zeros_like(self.params)
        self.state = torch.zeros(max(off_s, 4), dtype=torch.float32, device=device)
        any_l2 = any(s['l2'] for s in self.specs.values())
        self.l2 = torch.zeros_like(self.params) if any_l2 else None
        for name, s in self.specs.items():
            flat = self.params if s['trainable'] else self.state
            v = flat[s['offset']:s['offset'] + s['numel']].view(s['shape'])
            v.copy_(torch.from_numpy(s['init']))
            self._views[name] = v
            if s['trainable']:
                self._gviews[name] = self.grads[s['offset']:s['offset'] + s['numel']].view(s['shape'])
                if s['l2']:
                    self.l2[s['offset']:s['offset'] + s['numel']] = s['l2']
            s['init'] = None
        # Adam hyper-parameters and step counter live on the device (graph replay)
        self.version += 1
        self.hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-7], dtype=torch.float32, device=device)
        self.step = torch.zeros(1, dtype=torch.int64, device=device)

    def rebind_grads(self, full, reduced=None):
        """Moves the flat gradient buffer to other device memory (data parallelism over NVLink peer memory: the buffer
        must be an IPC-exportable allocation).  ``reduced``: where the cross-replica sums land (``get_grads`` / the loss
        scalar read from there); None = the buffer itself."""
        n = max(self.n_trainable, 4)
        assert full.numel() == n + 4 and full.dtype == torch.float32
        full.zero_()
        self.grads_full = full
        self.grads = full[:n]
        self.grads_reduced = reduced
        res = reduced if reduced is not None else full
        self.loss_slot = res[n:n + 1]
        self.loss_in = full[n:n + 1]       # the local term goes into the exchanged buffer, the cross-replica sum is read from `res`
        for name, s in self.specs.items():
            if s['trainable']:
                self._gviews[name] = self.grads[s['offset']:s['offset'] + s['numel']].view(s['shape'])
        self._rviews = None if reduced is None else {
            name: reduced[s['offset']:s['offset'] + s['numel']].view(s['shape']) for name, s in self.specs.items() if s['trainable']}

    def view(self, name):
        return self._views[name]

    def gview(self, name):
        return self._gviews[name]

    def ptr(self, name):
        return C.c_void_p(self._views[name].data_ptr()) if name in self._views else None

    def gptr(self, name):
        return C.c_void_p(self._gviews[name].data_ptr()) if name in self._gviews else None

    def names(self, trainable=None):
        return [k for k, s in self.specs.items() if trainable is None or s['trainable'] == trainable]

    def get_weights(self):
        out = OrderedDict()
        for name, s in self.specs.items():
            out[name] = self._views[name].detach().cpu().numpy().copy() if self.device is not None else s['init'].copy()
        return out

    def set_weights(self, weights):
        self.version += 1
        for name, arr in weights.items():
            if name not in self.specs:
                raise KeyError(f'unknown variable {name}')
            s = self.specs[name]
            arr = np.ascontiguousarray(arr, dtype=np.float32)
            if tuple(arr.shape) != tuple(s['shape']):
                raise ValueError(f'{name}: shape {arr.shape} != {s["shape"]}')
            if self.device is None:
                s['init'] = arr
            else:
                self._views[name].copy_(torch.from_numpy(arr))

    def get_grads(self):
        views = getattr(self, '_rviews', None) or self._gviews      # cross-replica sums when they live in their own buffer
        return OrderedDict((k, views[k].detach().cpu().numpy().copy()) for k in views)

    def count(self, trainable=None):
        return int(sum(int(np.prod(s['shape'])) for s in self.specs.values()
                       if trainable is None or s['trainable'] == trainable))


# ----------------------------------------------------------------------------
# buffers and views
# ----------------------------------------------------------------------------
class Buf:
    def __init__(self, plan, n, h, w, c, name, dtype=None, external=None, zero=False):
        self.plan, self.n, self.h, self.w, self.c, self.name = plan, n, h, w, c, name
        self.dtype = dtype or plan.dtype
        self.zero = zero
        self.data = external
        self.grad = None
        self.want_grad = False
        self.virtual = False          # output of a BatchNorm folded into its consumers: never materialised (gradient only)
        plan.bufs.append(self)

    def allocate(self, training):
        dev = self.plan.device
        if self.data is None and not self.virtual:
            alloc = torch.zeros if self.zero else torch.empty
            self.data = alloc(self.n, self.h, self.w, self.c, dtype=self.dtype, device=dev)
        if training and self.want_grad and self.grad is None:
            self.grad = torch.empty(self.n, self.h, self.w, self.c, dtype=self.dtype, device=dev)

    def nbytes(self):
        return self.n * self.h * self.w * self.c * (2 if self.dtype == torch.bfloat16 else 4)


class TRef:
    """Channels [coff, coff+c) of a Buf (= one logical tensor of the model)."""

    def __init__(self, buf: Buf, coff=0, c=None):
        self.buf, self.coff = buf, coff
        self.c = buf.c - coff if c is None else c
        self.act = None            # (code, alpha) when this tensor IS the output of conv+activation
        self.skip_consumed = False  # a second consumer (skip concat) adds into its gradient first
        self.needs_grad = True
        self.fold = None           # (pre-BN TRef, BNOp) when this tensor is the never-materialised output of a folded BatchNorm
        self._ct = self._gct = None

    n = property(lambda s: s.buf.n)
    h = property(lambda s: s.buf.h)
    w = property(lambda s: s.buf.w)
    shape = property(lambda s: (s.buf.n, s.buf.h, s.buf.w, s.c))

    def ct(self):
        if self._ct is None:
            assert self.buf.data is not None, f'{self.buf.name} is the output of a folded BatchNorm and has no storage'
            self._ct = N.tensor_view(self.buf.data, self.coff, self.c)
        return C.byref(self._ct)

    def src(self):
        """(tensor view to READ, [2C] scale|shift pointer or None): the pre-BN tensor and the BatchNorm's affine when this
        tensor is a folded BatchNorm output, else the tensor itself."""
        if self.fold is None:
            return self.ct(), None
        a, bn = self.fold
        return a.ct(), N.ptr(bn.ss)

    def gct(self):
        if self._gct is None:
            assert self.buf.grad is not None, f'no gradient buffer for {self.buf.name}'
            self._gct = N.tensor_view(self.buf.grad, self.coff, self.c)
        return C.byref(self._gct)

    def mask_args(self):
        """(mask view or None, act code, alpha) for kernels that write this tensor's gradient."""
        if self.act is None:
            return None, N.ACT_NONE, 0.0
        return self.ct(), self.act[0], self.act[1]

    def torch_view(self):
        return self.buf.data[..., self.coff:self.coff + self.c]

    def torch_grad(self):
        return self.buf.grad[..., self.coff:self.coff + self.c]


# ----------------------------------------------------------------------------
# ops
# ----------------------------------------------------------------------------
def conv_workspace(plan, taps, inputs, y, out_multiple=16):
    """bf16 weight-repack workspace of the tcgen05 kernels (dnnca_conv_workspace_bytes), or None when the layer
    cannot take the tensor-core path (fp32 mode, output channel counts that are not multiples of ``out_multiple``:
    16 in general, 8 for single-input Conv2D fprop, whose epilogue stores 8-channel groups)."""
    ins = [t for t in inputs if t is not None]
    if plan.dtype != torch.bfloat16 or y.c % out_multiple or y.coff % 8 or y.buf.c % 8 or any(t.buf.c % 8 or t.coff % 8 for t in ins):
        return None                       # (the library re-checks; narrow inputs only need 16-byte aligned pixels)
    if len(ins) > 1 and any(t.c % 16 for t in ins):
        return None
    nbytes = N.lib().dnnca_conv_workspace_bytes(taps, sum(t.c for t in ins), y.c)
    return torch.empty(nbytes, dtype=torch.uint8, device=plan.device)


def conv_workspace_ok(plan, inputs, y):
    """Would ``conv_workspace`` hand this layer a tensor-core workspace? (same conditions, no allocation)"""
    ins = [t for t in inputs if t is not None]
    if plan.dtype != torch.bfloat16 or y.c % 16 or y.coff % 8 or y.buf.c % 8 or any(t.buf.c % 8 or t.coff % 8 for t in ins):
        return False
    return not (len(ins) > 1 and any(t.c % 16 for t in ins))


def ws_args(ws):
    return (N.ptr(ws), ws.numel()) if ws is not None else (None, 0)


class Op:
    def fwd(self, train):
        raise NotImplementedError

    def bwd(self):
        pass

    def grad_params(self):
        """Names of the trainable variables whose gradients ``bwd`` writes (the data-parallel all-reduce launches a
        bucket of the flat gradient buffer as soon as every variable in it has been written)."""
        return ()


class ConvOp(Op):
    """layers.Conv2D (components.py:47-50,123-126; multiresunet.py:51-52).

    ``x2``: second input whose channels follow x's -- the conv reads both producers of
    ``tf.concat([tconv0, cropped], -1)`` (components.py:164), the concat is never materialised;
    backward writes the two gradients to their own tensors."""

    def __init__(self, plan, x: TRef, y: TRef, kernel, bias, ksize, act, stats=None, x2: TRef = None):
        self.p, self.x, self.x2, self.y, self.kernel, self.bias, self.k = plan, x, x2, y, kernel, bias, ksize
        self.act = act or (N.ACT_NONE, 0.0)
        self.stats = stats
        self.ws = None
        if act and act[0] != N.ACT_NONE:
            y.act = act

    def folded(self):
        return self.x.fold is not None or (self.x2 is not None and self.x2.fold is not None)

    def allocate(self, training):
        self.ws = conv_workspace(self.p, self.k * self.k, [self.x, self.x2], self.y) if self.ws is None else self.ws
        if self.folded() and getattr(self, 'scratch', None) is None:
            self.scratch = torch.empty(N.lib().dnnca_conv2d_fold_scratch_bytes(self.x.c + (self.x2.c if self.x2 else 0), self.y.c) // 4, dtype=torch.float32,
                                       device=self.p.device)

    def fwd(self, train):
        ps = self.p.params
        st = self.stats.fwd_ptr() if (self.stats and train) else None
        if self.folded():           # BatchNorm of the input(s) folded into this conv (bn_fold.cu)
            xa, fa = self.x.src()
            xb, fb = self.x2.src() if self.x2 else (None, None)
            N.call('dnnca_conv2d_fprop_affine', N.stream_ptr(), xa, xb, fa, fb, ps.ptr(self.kernel), ps.ptr(self.bias),
                   self.y.ct(), self.act[0], self.act[1], st, *ws_args(self.ws), N.ptr(self.scratch))
            return
        N.call('dnnca_conv2d_fprop', N.stream_ptr(), self.x.ct(), self.x2.ct() if self.x2 else None,
               ps.ptr(self.kernel), ps.ptr(self.bias), self.y.ct(), self.k, self.act[0], self.act[1], st, *ws_args(self.ws))

    def grad_params(self):
        return tuple(n for n in (self.kernel, self.bias) if n)

    def bwd(self):
        self.bwd_params()
        self.bwd_input()

    def bwd_params(self):
        """wgrad only: reads x and dz, writes the parameter gradients -- nothing downstream in the backward pass depends
        on it, so ``Plan.backward`` may issue it on a side stream beside the dgrad chain."""
        ps = self.p.params
        if self.folded():
            xa, fa = self.x.src()
            xb, fb = self.x2.src() if self.x2 else (None, None)
            N.call('dnnca_conv2d_wgrad_affine', N.stream_ptr(), xa, xb, fa, fb, self.y.gct(), ps.gptr(self.kernel),
                   ps.gptr(self.bias), N.ptr(self.scratch))
        else:
            N.call('dnnca_conv2d_wgrad', N.stream_ptr(), self.x.ct(), self.x2.ct() if self.x2 else None, self.y.gct(),
                   ps.gptr(self.kernel), ps.gptr(self.bias), self.k)

    def bwd_input(self):
        """dgrad only (also the input-gradient chain of callbacks.py:290-299, which needs no parameter gradients)."""
        if self.x.needs_grad:
            ps = self.p.params
            bn = getattr(self.x, 'bnr', None)
            if bn is not None and self.p.train_bn:     # x is a BatchNorm output read by this conv alone: its backward sums
                N.call('dnnca_conv2d_dgrad_bnreduce', N.stream_ptr(), self.y.gct(), ps.ptr(self.kernel), self.x.gct(),
                       self.x2.gct() if self.x2 else None, self.k, bn.x.ct(), N.ptr(bn.mi), bn.stats.bwd_ptr(),
                       *ws_args(self.ws))
                return
            m, a, al = self.x.mask_args()
            N.call('dnnca_conv2d_dgrad', N.stream_ptr(), self.y.gct(), ps.ptr(self.kernel), self.x.gct(),
                   self.x2.gct() if self.x2 else None, self.k, m, a, al, *ws_args(self.ws))


class TConvOp(Op):
    """layers.Convolution2DTranspose k=s=2 (components.py:118-120; multiresunet.py:200-215)."""

    def __init__(self, plan, x, y, kernel, bias, stats=None):
        self.p, self.x, self.y, self.kernel, self.bias, self.stats = plan, x, y, kernel, bias, stats
        self.ws = None

    def allocate(self, training):
        self.ws = conv_workspace(self.p, 4, [self.x], self.y) if self.ws is None else self.ws

    def fwd(self, train):
        ps = self.p.params
        N.call('dnnca_convtranspose2x2_fprop', N.stream_ptr(), self.x.ct(), ps.ptr(self.kernel), ps.ptr(self.bias),
               self.y.ct(), self.stats.fwd_ptr() if (self.stats and train) else None, *ws_args(self.ws))

    def grad_params(self):
        return tuple(n for n in (self.kernel, self.bias) if n)

    def bwd(self):
        self.bwd_params()
        self.bwd_input()

    def bwd_params(self):
        ps = self.p.params
        N.call('dnnca_convtranspose2x2_wgrad', N.stream_ptr(), self.x.ct(), self.y.gct(), ps.gptr(self.kernel),
               ps.gptr(self.bias))

    def bwd_input(self):
        if self.x.needs_grad:
            ps = self.p.params
            bn = getattr(self.x, 'bnr', None)
            if bn is not None and self.p.train_bn:
                N.call('dnnca_convtranspose2x2_dgrad_bnreduce', N.stream_ptr(), self.y.gct(), ps.ptr(self.kernel), self.x.gct(),
                       bn.x.ct(), N.ptr(bn.mi), bn.stats.bwd_ptr(), *ws_args(self.ws))
                return
            m, a, al = self.x.mask_args()
            N.call('dnnca_convtranspose2x2_dgrad', N.stream_ptr(), self.y.gct(), ps.ptr(self.kernel), self.x.gct(), m, a,
                   al, *ws_args(self.ws))


class PoolOp(Op):
    """layers.MaxPool2D([2,2], strides=2) (components.py:54)."""

    def __init__(self, plan, x, y, stats=None):
        self.p, self.x, self.y, self.stats = plan, x, y, stats
        self.idx = None

    def allocate(self, training):
        if training and self.idx is None:
            self.idx = torch.empty(self.y.n, self.y.h, self.y.w, self.y.c, dtype=torch.uint8, device=self.p.device)

    def fwd(self, train):
        keep = train or self.p.want_input_grad       # the input-gradient chain runs an inference forward but needs the argmax
        st = self.stats.fwd_ptr() if (self.stats and train) else None
        if self.x.fold is not None:
            xa, fa = self.x.src()
            N.call('dnnca_maxpool2x2_fwd_affine', N.stream_ptr(), xa, fa, self.y.ct(), N.ptr(self.idx) if keep else None, st)
            return
        N.call('dnnca_maxpool2x2_fwd', N.stream_ptr(), self.x.ct(), self.y.ct(), N.ptr(self.idx) if keep else None, st)

    def bwd(self):
        if not self.x.needs_grad:
            return
        m, a, al = self.x.mask_args()
        dskip = self.x.gct() if self.x.skip_consumed else None
        N.call('dnnca_maxpool2x2_bwd', N.stream_ptr(), self.y.gct(), N.ptr(self.idx), dskip, self.x.gct(), m, a, al)

    def bwd_input(self):
        self.bwd()


class BNStats:
    """fp64 accumulators of one BatchNormalization: forward (sum, sumsq) and backward (sum dy, sum dy*xhat)."""

    def __init__(self, plan, c):
        self.p, self.c = plan, c
        self.off = plan.stats_len
        plan.stats_len += 4 * c

    def fwd_ptr(self):
        return C.c_void_p(self.p.stats.data_ptr() + 8 * self.off)

    def bwd_ptr(self):
        return C.c_void_p(self.p.stats.data_ptr() + 8 * (self.off + 2 * self.c))


class BNOp(Op):
    """layers.BatchNormalization (components.py:57,59,130,131).  Training statistics arrive
    in ``stats`` (filled by the producer's epilogue or a channel_stats pass)."""

    def __init__(self, plan, x, y, prefix, stats: BNStats, scale=True, fused_stats=True):
        self.p, self.x, self.y, self.prefix, self.stats = plan, x, y, prefix, stats
        self.gamma = f'{prefix}/gamma' if scale else None
        self.fused_stats = fused_stats
        self.folded = False        # Plan.fold_batchnorms: the consumers read `x` + the affine, `y` is never written
        self.reduce_fused = False  # Plan.fuse_bn_reductions: the sole consumer's dgrad takes the backward sums
        self.ss = self.mi = None   # scale|shift and mean|invstd, fp32 [2C] each

    def allocate(self, training):
        if self.ss is None:
            self.ss = torch.empty(2 * self.x.c, dtype=torch.float32, device=self.p.device)
            self.mi = torch.empty(2 * self.x.c, dtype=torch.float32, device=self.p.device)

    def fwd(self, train):
        ps, s, c = self.p.params, N.stream_ptr(), self.x.c
        g = ps.ptr(self.gamma) if self.gamma else None
        if train:
            if not self.fused_stats:
                N.call('dnnca_channel_stats', s, self.x.ct(), self.stats.fwd_ptr())
            N.call('dnnca_bn_finalize', s, self.stats.fwd_ptr(), self.x.n * self.x.h * self.x.w, c, g,
                   ps.ptr(f'{self.prefix}/beta'), BN_MOMENTUM, BN_EPSILON, ps.ptr(f'{self.prefix}/moving_mean'),
                   ps.ptr(f'{self.prefix}/moving_var'), N.ptr(self.ss), N.ptr(self.mi))
        else:
            N.call('dnnca_bn_inference_params', s, c, g, ps.ptr(f'{self.prefix}/beta'), BN_EPSILON,
                   ps.ptr(f'{self.prefix}/moving_mean'), ps.ptr(f'{self.prefix}/moving_var'), N.ptr(self.ss))
        if not self.folded:
            N.call('dnnca_bn_apply', s, self.x.ct(), N.ptr(self.ss), self.y.ct())

    def grad_params(self):
        return tuple(n for n in (self.gamma, f'{self.prefix}/beta') if n)

    def bwd(self):
        ps, s = self.p.params, N.stream_ptr()
        if not self.reduce_fused:      # else the dgrad that wrote y's gradient left the sums (dnnca_*_dgrad_bnreduce)
            N.call('dnnca_bn_bwd_reduce', s, self.x.ct(), self.y.gct(), N.ptr(self.mi), self.stats.bwd_ptr())
        act = self.x.act or (N.ACT_NONE, 0.0)
        N.call('dnnca_bn_bwd_apply', s, self.x.ct(), self.y.gct(), N.ptr(self.mi),
               ps.ptr(self.gamma) if self.gamma else None, self.stats.bwd_ptr(), self.x.gct(), act[0], act[1],
               ps.gptr(self.gamma) if self.gamma else None, ps.gptr(f'{self.prefix}/beta'))


    def bwd_input(self):
        """Backward of an INFERENCE-mode BatchNormalization (moving statistics are constants): dx = dy * scale
        (* act'(x)) -- ``dnnca_bn_bwd_apply`` with zero batch sums and invstd := the folded scale."""
        c = self.x.c
        if getattr(self, 'mi_inf', None) is None:
            self.mi_inf = torch.zeros(2 * c, dtype=torch.float32, device=self.p.device)
            self.zero_sums = torch.zeros(2 * c, dtype=torch.float64, device=self.p.device)
        self.mi_inf[c:].copy_(self.ss[:c])
        act = self.x.act or (N.ACT_NONE, 0.0)
        N.call('dnnca_bn_bwd_apply', N.stream_ptr(), self.x.ct(), self.y.gct(), N.ptr(self.mi_inf), None,
               N.ptr(self.zero_sums), self.x.gct(), act[0], act[1], None, None)


class ConvertOp(Op):
    """fp32 network input -> activation dtype (the tail of data.py:193-206 on the device)."""

    def __init__(self, plan, x, y):
        self.plan, self.x, self.y = plan, x, y
        if y.coff == 0 and y.c == y.buf.c and x.coff == 0 and x.c == x.buf.c and x.c == y.c:
            plan.input_cast = y          # dense cast target: uint8 batches can be staged straight into it (Model._load_batch)

    def fwd(self, train):
        if getattr(self.plan, 'prestaged', False):
            return                       # the batch was written in the activation dtype by dnnca_u8_to_unit
        N.call('dnnca_convert', N.stream_ptr(), self.x.ct(), self.y.ct())

    def bwd_input(self):
        if self.y.needs_grad and self.x.buf.grad is not None:       # gradient w.r.t. the fp32 network input
            N.call('dnnca_convert', N.stream_ptr(), self.y.gct(), self.x.gct())


class AddReluAffineOp(Op):
    """BN -> add -> relu -> BN tail of MultiResBlock / ResPath (multiresunet.py:120-124,148-150), inference."""

    def __init__(self, plan, a, fa, b, fb, fo, y):
        self.p, self.a, self.fa, self.b, self.fb, self.fo, self.y = plan, a, fa, b, fb, fo, y

    def fwd(self, train):
        assert not train, 'MultiResUnet is forward/inference only in this build'
        N.call('dnnca_add_relu_affine', N.stream_ptr(), self.a.ct(), N.ptr(self.fa() if self.fa else None),
               self.b.ct(), N.ptr(self.fb() if self.fb else None), N.ptr(self.fo() if self.fo else None), self.y.ct())


class CallbackOp(Op):
    """Host-side closure run at this point of the forward sequence (e.g. folding BN into weights)."""

    def __init__(self, fn):
        self.fn = fn

    def fwd(self, train):
        self.fn(train)


# ----------------------------------------------------------------------------
# plan
# ----------------------------------------------------------------------------
class Plan:
    def __init__(self, params: ParamStore, batch, height, width, channels, dtype, device, want_input_grad=False):
        self.params, self.dtype, self.device = params, dtype, device
        self.want_input_grad = bool(want_input_grad)     # callbacks.py:290-299: d(output)/d(input) on request
        self.bufs: list[Buf] = []
        self.ops: list[Op] = []
        self.stats_len = 0
        self.stats = None
        self.batch, self.height, self.width, self.channels = batch, height, width, channels
        # static fp32 input / label / output buffers (the H2D copies of a step land here)
        self.x_in = torch.zeros(batch, height, width, channels, dtype=torch.float32, device=device)
        self.y_in = torch.zeros(batch, height, width, dtype=torch.float32, device=device)
        self.input = TRef(Buf(self, batch, height, width, channels, 'input', torch.float32, external=self.x_in))
        self.input.needs_grad = self.want_input_grad
        self.features = None          # TRef feeding the head
        self.head = None              # (kernel name, bias name)
        self.allocated_training = None
        self.graphs = {}
        self.train_bn = True          # BatchNorm layers run in training mode in the backward pass being launched
        self.branches = []            # [(first op, end op)] of mutually independent, adjacent op ranges (set by the model)

    def new_buf(self, h, w, c, name, n=None, zero=False):
        return Buf(self, n or self.batch, h, w, c, name, zero=zero)

    def add(self, op):
        self.ops.append(op)
        return op

    def fold_batchnorms(self, min_channels=32):
        """Folds every BatchNorm whose output is read only by 3x3 convs the folded tensor-core kernel serves and by
        max-pools into those consumers (bn_fold.cu): the BN's apply pass and its output tensor disappear, the consumers
        read the pre-BN tensor + the [2C] scale|shift the BN's finalize wrote.  A BN feeding the head, a transposed conv,
        a concat slice (MulmoUNet's bottleneck) or a layer narrower than ``min_channels`` (those run on the
        row-Toeplitz kernels) keeps its materialised output.  Returns the number of folded BatchNorms."""
        if self.dtype != torch.bfloat16 or self.allocated_training is not None:
            return 0
        lib = N.lib()
        n = 0
        for bn in [op for op in self.ops if isinstance(op, BNOp)]:
            t, a = bn.y, bn.x
            if t is self.features or t.coff != 0 or t.c != t.buf.c or a.c < min_channels or t.c % 8:
                continue
            users = [op for op in self.ops if getattr(op, 'x', None) is t or getattr(op, 'x2', None) is t]
            if not users or any(op is bn for op in users):
                continue
            ok = True
            for op in users:
                if isinstance(op, PoolOp):
                    continue
                if not (isinstance(op, ConvOp) and op.k == 3):
                    ok = False
                    break
                # shape query with stand-in views (no storage yet): same n,h,w,c / strides as the real tensors
                def view(tr):
                    if tr is None:
                        return None
                    src = tr.fold[0] if tr.fold else (a if tr is t else tr)
                    v = N.Tensor(256, src.n, src.h, src.w, src.c, src.buf.c, src.coff, N.BF16)
                    return C.byref(v)
                if not lib.dnnca_conv2d_fold_supported(view(op.x), view(op.x2), view(op.y), 3) or \
                        conv_workspace_ok(self, [op.x, op.x2], op.y) is False:
                    ok = False
                    break
            if not ok:
                continue
            t.fold = (a, bn)
            t.buf.virtual = True
            bn.folded = True
            n += 1
        self.n_folded_bn = n
        return n

    def fuse_bn_reductions(self):
        """BatchNorm backward sums taken by the dgrad that writes the BN output's gradient: applies to every BatchNorm
        whose output has exactly ONE reader, a Conv2D (as its first input) or a transposed conv -- then that reader's
        dgrad is the only writer of the gradient and `bn_bwd_reduce` (a full read of x and dy) is dropped.  Skip tensors
        (read by the pool and by the decoder) and the BN feeding the head keep the separate pass."""
        import os
        # OFF by default: measured on B200 (profiles/r02p_*) it does not pay -- unet_big B=32 12.396 -> 12.383 ms, mulmo_unet
        # 5.64 -> 5.74 ms.  The 64-channel dgrads at 256^2 already move 3.7 TB/s; reading the BN input as well makes them
        # HBM-bound, so the pass only moves from one kernel into another.  DNNCA_BN_REDUCE_FUSE=1 enables it.
        if os.environ.get('DNNCA_BN_REDUCE_FUSE', '0') != '1' or self.want_input_grad:
            return 0
        n = 0
        for bn in [op for op in self.ops if isinstance(op, BNOp)]:
            t = bn.y
            if t is self.features:
                continue
            users = [op for op in self.ops if op is not bn and (getattr(op, 'x', None) is t or getattr(op, 'x2', None) is t)]
            if len(users) != 1 or not isinstance(users[0], (ConvOp, TConvOp)) or users[0].x is not t:
                continue
            if t.coff != 0 or t.c != t.buf.c or bn.x.c != t.c:
                continue
            t.bnr = bn
            bn.reduce_fused = True
            n += 1
        self.n_fused_bn_reduce = n
        return n

    def allocate(self, training):
        if self.allocated_training is not None and (self.allocated_training or not training):
            return
        for op in self.ops:
            for t in (getattr(op, 'x', None), getattr(op, 'x2', None), getattr(op, 'y', None)):
                if training and isinstance(t, TRef) and t.needs_grad:
                    t.buf.want_grad = True
        if training and self.features is not None:
            self.features.buf.want_grad = True
        for b in self.bufs:
            b.allocate(training)
        for op in self.ops:
            if hasattr(op, 'allocate'):
                op.allocate(training)
        if self.stats is None:
            self.stats = torch.zeros(max(self.stats_len, 1), dtype=torch.float64, device=self.device)
        B, H, W = self.batch, self.height, self.width
        if getattr(self, 'logits', None) is None:
            self.logits = torch.empty(B, H, W, 1, dtype=torch.float32, device=self.device)
            self.probs = torch.empty(B, H, W, 1, dtype=torch.float32, device=self.device)
        if training and getattr(self, 'per_sample', None) is None:
            self.per_sample = torch.zeros(B, dtype=torch.float32, device=self.device)
            self.lstats = torch.zeros(16, dtype=torch.uint8, device=self.device)
        self.allocated_training = training or bool(self.allocated_training)

    def activation_bytes(self):
        """Bytes of activation + activation-gradient storage actually allocated by this plan."""
        return sum(t.numel() * t.element_size() for b in self.bufs for t in (b.data, b.grad) if t is not None)

    # ---- launch sequences ----------------------------------------------------
    def _branch_streams(self):
        """One extra stream per independent branch beyond the first (MulmoUNet's per-modality encoders, unet.py:182-186)
        or None when the plan has no branches / DNNCA_BRANCH_STREAMS=0."""
        import os
        if len(self.branches) < 2 or os.environ.get('DNNCA_BRANCH_STREAMS', '1') == '0':
            return None
        if getattr(self, '_bstreams', None) is None:
            self._bstreams = [torch.cuda.Stream(device=self.device) for _ in self.branches[1:]]
            self._bsides = [torch.cuda.Stream(device=self.device) for _ in self.branches[1:]]
            for (s0, e0), (s1, e1) in zip(self.branches[:-1], self.branches[1:]):
                assert e0 == s1, 'branches must be adjacent op ranges'
        return self._bstreams

    def forward(self, train=False):
        bs = self._branch_streams()
        if bs is None:
            for op in self.ops:
                op.fwd(train)
            return
        # independent branches run side by side in the captured graph: branch 0 stays on the launching stream, the
        # others fork from an event recorded after the ops in front of them and join before the first op behind them
        first, last = self.branches[0][0], self.branches[-1][1]
        main = torch.cuda.current_stream()
        for op in self.ops[:first]:
            op.fwd(train)
        fork = torch.cuda.Event()
        fork.record(main)
        joins = []
        for k, (s, e) in enumerate(self.branches):
            if k == 0:
                for op in self.ops[s:e]:
                    op.fwd(train)
                continue
            st = bs[k - 1]
            st.wait_event(fork)
            with torch.cuda.stream(st):
                for op in self.ops[s:e]:
                    op.fwd(train)
                ev = torch.cuda.Event()
                ev.record(st)
            joins.append(ev)
        for ev in joins:
            main.wait_event(ev)
        for op in self.ops[last:]:
            op.fwd(train)

    def head_forward(self):
        ps = self.params
        N.call('dnnca_head_fwd', N.stream_ptr(), self.features.ct(), ps.ptr(self.head[0]), ps.ptr(self.head[1]),
               N.ptr(self.logits), N.ptr(self.probs))

    def head_loss(self, loss_cfg: N.LossConfig, with_grads=True):
        ps, s = self.params, N.stream_ptr()
        if getattr(self, 'per_sample', None) is None:
            self.per_sample = torch.zeros(self.batch, dtype=torch.float32, device=self.device)
            self.lstats = torch.zeros(16, dtype=torch.uint8, device=self.device)
        self.per_sample.zero_()
        if not loss_cfg.has_weight:
            N.call('dnnca_label_stats_init', s, N.ptr(self.lstats))
            N.call('dnnca_label_stats', s, N.ptr(self.y_in), self.y_in.numel(), N.ptr(self.lstats))
        f = self.features
        act = f.act or (N.ACT_NONE, 0.0)
        if self.head is None:       # inference-only model whose head weights are folded (MultiResUnet conv10 + BN)
            assert not with_grads, 'a folded head has no gradient path'
            wf, bf = self._head_fold[0], self._head_fold[1]
            N.call('dnnca_head_bce_fwd_bwd', s, f.ct(), N.ptr(wf), N.ptr(bf), N.ptr(self.y_in), N.ptr(self.lstats),
                   C.byref(loss_cfg), N.ptr(self.logits), N.ptr(self.probs), N.ptr(self.per_sample), None, act[0], act[1],
                   None, None)
            return
        N.call('dnnca_head_bce_fwd_bwd', s, f.ct(), ps.ptr(self.head[0]), ps.ptr(self.head[1]), N.ptr(self.y_in),
               N.ptr(self.lstats), C.byref(loss_cfg), N.ptr(self.logits), N.ptr(self.probs), N.ptr(self.per_sample),
               f.gct() if with_grads else None, act[0], act[1], ps.gptr(self.head[0]) if with_grads else None,
               ps.gptr(self.head[1]) if with_grads else None)

    def head_input_grad(self):
        f = self.features
        act = f.act or (N.ACT_NONE, 0.0)
        N.call('dnnca_head_input_grad', N.stream_ptr(), f.ct(), self.params.ptr(self.head[0]), self.params.ptr(self.head[1]),
               f.gct(), act[0], act[1])

    def backward_inputs(self):
        """dgrad chain only, down to the network input (no parameter gradients are touched)."""
        for op in reversed(self.ops):
            if hasattr(op, 'bwd_input'):
                op.bwd_input()

    def input_grad_f32(self):
        g = self.input.buf.grad
        if g is None:
            raise RuntimeError('this plan was not built with want_input_grad=True')
        return g

    def _bwd_range(self, ops, side_stream, main, after_op=None, base=0):
        """Backward of ``ops`` (already reversed) on stream ``main`` (the current one); weight gradients go to
        ``side_stream`` when given (noted in ``_open_sides`` until the next join)."""
        for i, op in enumerate(ops):
            if side_stream is not None and hasattr(op, 'bwd_params'):
                ev = torch.cuda.Event()
                ev.record(main)
                side_stream.wait_event(ev)
                with torch.cuda.stream(side_stream):
                    op.bwd_params()
                if side_stream not in self._open_sides:
                    self._open_sides.append(side_stream)
                op.bwd_input()
            else:
                op.bwd()
            if after_op is not None:
                after_op(base + i)

    def backward(self, after_op=None, side_stream=None):
        """The tape: forward list reversed.  ``after_op(i)`` is called after the i-th backward op was launched
        (data parallelism: launches the gradient buckets that just became complete; the callback joins the side
        streams itself before a bucket leaves, see ``join_side_streams``).

        ``side_stream``: the weight-gradient kernels are issued there (fork on an event recorded when dz is complete,
        join at the end), so that in the captured graph a wgrad is a sibling of the dgrad chain instead of a link in
        it: its prologue and its last, partly filled wave overlap the next dgrad.  Independent branches
        (``self.branches``) run their backward chains on their own streams."""
        main = torch.cuda.current_stream()
        self._open_sides = []
        bs = self._branch_streams() if side_stream is not None else None
        rev = list(reversed(self.ops))
        n = len(rev)
        if bs is None:
            self._bwd_range(rev, side_stream, main, after_op)
            self.join_side_streams()
            return
        first, last = self.branches[0][0], self.branches[-1][1]
        self._bwd_range(rev[:n - last], side_stream, main, after_op)          # everything behind the branches (decoder)
        fork = torch.cuda.Event()
        fork.record(main)
        # branches in reverse order: the last branch's ops come first on the tape; branch 0 stays on the launching stream
        for k in range(len(self.branches) - 1, -1, -1):
            s, e = self.branches[k]
            ops = rev[n - e:n - s]
            if k == 0:
                # (every other branch has been issued by now: a bucket leaving from here joins them first)
                self._bwd_range(ops, side_stream, main, after_op, base=n - e)
                continue
            st, sd = bs[k - 1], self._bsides[k - 1]
            st.wait_event(fork)
            with torch.cuda.stream(st):
                self._bwd_range(ops, sd, st, None)
            self._open_sides.append(st)
        self._bwd_range(rev[n - first:], None, main, after_op, base=n - first)
        self.join_side_streams()

    def join_side_streams(self):
        """The launching stream waits for everything issued so far on the side streams of this backward pass."""
        main = torch.cuda.current_stream()
        for st in getattr(self, '_open_sides', []):
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
        self._open_sides = []

    def ready_frontier(self):
        """ready[i] = lowest offset R of the flat gradient buffer such that every gradient in [R, end) has been
        written once backward op i has run: the highest end (offset + numel) over the variables still to be written
        by later backward ops (earlier layers sit at lower offsets, so R falls as the backward pass proceeds)."""
        specs = self.params.specs
        ends = []
        for op in reversed(self.ops):
            e = 0
            for n in op.grad_params():
                sp = specs[n]
                if sp['trainable']:
                    e = max(e, sp['offset'] + sp['numel'])
            ends.append(e)
        ready, pending = [0] * len(ends), 0
        for i in range(len(ends) - 1, -1, -1):
            ready[i] = pending
            pending = max(pending, ends[i])
        self.pending_before_backward = pending      # frontier before the first backward op (head gradients are done)
        return ready

    def zero_step_state(self):
        self.params.grads_full.zero_()
        if self.stats_len:
            self.stats.zero_()

    def end_backward(self):
        """Hook of the training launch sequences, run after the backward pass has been issued and its side streams joined
        (plans that compute on padded copies of the variables hand the gradients back here)."""
