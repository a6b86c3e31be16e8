"""torch-CPU restatement of the reference's model graphs and training step.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``); parity unpinned.

Follows ``annotator/models/tf_models/components.py:16-320`` (block structure,
filter progression, skip order), ``unet.py:19-300`` and
``multiresunet.py:31-223``; the loss wiring of ``engine.py:270-286`` /
``losses.py:40-84``.  ``torch.autograd`` stands in for ``tf.GradientTape``.

Weight naming (own convention; the reference relies on Keras auto-names):
  U-Net family   ``enc[/m]/d{i}/conv{k}/{kernel,bias}``, ``.../bn{k}/{gamma,beta,moving_mean,moving_var}``,
                 ``.../pool_bn/...``, ``dec/u{j}/tconv/{kernel,bias}``, ``dec/u{j}/tconv_bn/...``,
                 ``dec/u{j}/conv{k}/...``, ``dec/u{j}/bn{k}/...``, ``head/{kernel,bias}``
  MultiResUnet   ``conv{n}/kernel``, ``bn{n}/...``, ``tconv{n}/{kernel,bias}`` numbered in the
                 reference's layer-creation order.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from . import ref_ops as ops


# ----------------------------------------------------------------------------
# initialisers  [TF-semantics] keras defaults: glorot_uniform kernels, zero bias,
# BN gamma=1 beta=0 moving_mean=0 moving_var=1
# ----------------------------------------------------------------------------
def glorot_uniform(rng: np.random.Generator, shape):
    """[TF-semantics] ``VarianceScaling(1.0,'fan_avg','uniform')``: for a 4-D
    kernel fan_in = shape[-2]*kh*kw, fan_out = shape[-1]*kh*kw; limit = sqrt(6/(fan_in+fan_out))."""
    rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


def solve_activation(identifier):
    """components.py:323-335 -> the tokens ``ref_ops.activation`` understands."""
    if identifier is None:
        return None
    if isinstance(identifier, str):
        return identifier
    if isinstance(identifier, dict):
        if identifier.get('class_name') == 'LeakyReLU':
            return ('leaky', float(identifier.get('config', {}).get('alpha', 0.3)))
        if identifier.get('class_name') == 'ReLU':
            return 'relu'
    if isinstance(identifier, tuple):
        return identifier
    raise ValueError(f'Failed to resolve activation: {identifier}')


def solve_regularizer(identifier):
    """kernel_regularizer.yaml:1-4 -> l2 coefficient (None = no regulariser)."""
    if identifier is None:
        return None
    if isinstance(identifier, dict) and identifier.get('class_name') in ('L2', 'l2'):
        return float(identifier.get('config', {}).get('l2', 0.01))
    if isinstance(identifier, str) and identifier.lower() == 'l2':
        return 0.01
    raise ValueError(f'unsupported kernel_regularizer {identifier}')


class _Base:
    def __init__(self):
        self.weights: 'OrderedDict[str, torch.Tensor]' = OrderedDict()
        self.trainable: list[str] = []
        self.regularized: list[str] = []
        self.l2 = None
        self.dtype = torch.float32
        self.device = torch.device('cpu')

    def to(self, device):
        """Evaluate the SAME restatement with torch on another device (tests/tools/parity_real_shapes.py runs the fp32
        oracle at the configs' full sizes on the GPU box's device with TF32 disabled; checked against the CPU
        evaluation at a small size there).  Test infrastructure only."""
        self.device = torch.device(device)
        for k in self.weights:
            self.weights[k] = self.weights[k].to(self.device)
        return self

    # -- variable creation ----------------------------------------------------
    def _add(self, name, array, trainable=True, regularized=False):
        assert name not in self.weights, name
        self.weights[name] = torch.tensor(np.asarray(array), dtype=self.dtype, device=self.device)
        if trainable:
            self.trainable.append(name)
        if regularized and self.l2 is not None:
            self.regularized.append(name)

    def _add_conv(self, rng, prefix, kh, kw, cin, cout, bias=True):
        self._add(f'{prefix}/kernel', glorot_uniform(rng, (kh, kw, cin, cout)), regularized=True)
        if bias:
            self._add(f'{prefix}/bias', np.zeros(cout, np.float32))

    def _add_tconv(self, rng, prefix, k, cin, cout):
        self._add(f'{prefix}/kernel', glorot_uniform(rng, (k, k, cout, cin)), regularized=True)
        self._add(f'{prefix}/bias', np.zeros(cout, np.float32))

    def _add_bn(self, prefix, c, scale=True):
        if scale:
            self._add(f'{prefix}/gamma', np.ones(c, np.float32))
        self._add(f'{prefix}/beta', np.zeros(c, np.float32))
        self._add(f'{prefix}/moving_mean', np.zeros(c, np.float32), trainable=False)
        self._add(f'{prefix}/moving_var', np.ones(c, np.float32), trainable=False)

    def _bn(self, ctx, prefix, x):
        w = ctx['w']
        y, mm, mv = ops.batchnorm(
            x, w.get(f'{prefix}/gamma'), w[f'{prefix}/beta'],
            w[f'{prefix}/moving_mean'], w[f'{prefix}/moving_var'], ctx['training'])
        if ctx['training']:
            ctx['new_moving'][f'{prefix}/moving_mean'] = mm
            ctx['new_moving'][f'{prefix}/moving_var'] = mv
        return y

    # -- public ---------------------------------------------------------------
    def get_weights(self):
        return OrderedDict((k, v.detach().cpu().numpy().copy()) for k, v in self.weights.items())

    def set_weights(self, weights):
        for k, v in weights.items():
            assert k in self.weights, k
            assert tuple(self.weights[k].shape) == tuple(np.shape(v)), (k, self.weights[k].shape, np.shape(v))
            self.weights[k] = torch.tensor(np.asarray(v), dtype=self.dtype, device=self.device)

    def randomize_bn(self, seed=1):
        """Non-trivial BN parameters for tests (gamma~U[.5,1.5], beta~N(0,.1))."""
        rng = np.random.default_rng(seed)
        for k in self.weights:
            c = self.weights[k].shape[0]
            if k.endswith('/gamma'):
                self.weights[k] = torch.tensor(rng.uniform(0.5, 1.5, c), dtype=self.dtype, device=self.device)
            elif k.endswith('/beta'):
                self.weights[k] = torch.tensor(rng.normal(0, 0.1, c), dtype=self.dtype, device=self.device)
            elif k.endswith('/moving_mean'):
                self.weights[k] = torch.tensor(rng.normal(0, 0.1, c), dtype=self.dtype, device=self.device)
            elif k.endswith('/moving_var'):
                self.weights[k] = torch.tensor(rng.uniform(0.5, 1.5, c), dtype=self.dtype, device=self.device)
            elif k.endswith('/bias'):
                self.weights[k] = torch.tensor(rng.normal(0, 0.05, c), dtype=self.dtype, device=self.device)

    def __call__(self, x, training=False):
        return self.forward(x, training=training)['probs']

    def forward(self, x, training=False, weights=None):
        ctx = dict(w=self.weights if weights is None else weights, training=training,
                   new_moving=OrderedDict(), pool_idx=[], tensors=OrderedDict())
        x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(device=self.device, dtype=self.dtype)
        logits = self._graph(ctx, x)
        return dict(logits=logits, probs=torch.sigmoid(logits), new_moving=ctx['new_moving'],
                    pool_idx=ctx['pool_idx'], tensors=ctx['tensors'])

    def train_step_grads(self, x, y, loss_config=None, n_replicas=1):
        """One ``Model.train_step`` up to (not including) the optimizer
        [TF-semantics]: loss = mean_B(per-sample loss) + sum(regulariser losses),
        both divided by the replica count under a MirroredStrategy; gradients
        w.r.t. trainable variables only."""
        loss_config = dict(loss_config or {})
        w = OrderedDict((k, v.clone().requires_grad_(k in self.trainable)) for k, v in self.weights.items())
        out = self.forward(x, training=True, weights=w)
        y = torch.as_tensor(np.asarray(y) if not torch.is_tensor(y) else y).to(device=self.device, dtype=self.dtype)
        if loss_config.pop('label_smoothing', False):
            y = ops.gaussian_filter2d(y, loss_config.pop('label_smoothing_filter_size', 6),
                                      loss_config.pop('label_smoothing_sigma', 3))
        loss_config.pop('label_smoothing_filter_size', None)
        loss_config.pop('label_smoothing_sigma', None)
        per_sample = ops.weighted_crossentropy(y, out['logits'], **loss_config)
        data_loss = per_sample.mean()
        reg = ops.l2_regularizer([w[k] for k in self.regularized], self.l2) if self.regularized else 0.0
        total = (data_loss + reg) / n_replicas
        grads = torch.autograd.grad(total, [w[k] for k in self.trainable], allow_unused=True)
        gd = OrderedDict()
        for k, g in zip(self.trainable, grads):
            gd[k] = torch.zeros_like(w[k]) if g is None else g.detach()
        return dict(loss=float((data_loss + reg).detach()), data_loss=float(data_loss.detach()), per_sample=per_sample.detach(),
                    grads=gd, logits=out['logits'].detach(), probs=out['probs'].detach(),
                    new_moving=out['new_moving'], pool_idx=out['pool_idx'])

    def n_params(self, trainable_only=False):
        names = self.trainable if trainable_only else list(self.weights)
        return int(sum(self.weights[k].numel() for k in names))


class RefUNetAnnotator(_Base):
    """``UNetAnnotator`` (unet.py:194-282) = ``UNet`` (unet.py:19-88) + 1x1 sigmoid head;
    with ``mulmo=True``: ``MulmoUNetAnnotator`` (unet.py:91-191,285-300)."""

    def __init__(self, n_filters_first, n_downsample, rate, kernel_size, conv_stride,
                 bn=False, padding='valid', activation='relu', kernel_regularizer=None,
                 mulmo=False, reference_index=0, n_conv=2):
        super().__init__()
        assert conv_stride == 1, 'reference configs use conv_stride 1 only'
        self.F, self.n, self.rate, self.k = n_filters_first, n_downsample, rate, kernel_size
        self.bn, self.padding = bn, padding
        self.act = solve_activation(activation)
        self.l2 = solve_regularizer(kernel_regularizer)
        self.mulmo, self.reference_index, self.n_conv = mulmo, reference_index, n_conv

    def build(self, input_shape, seed=0, dtype=torch.float32):
        self.dtype = dtype
        rng = np.random.default_rng(seed)
        cin_total = input_shape[-1]
        self.n_enc = cin_total if self.mulmo else 1          # unet.py:152-165
        enc_names = [f'enc/{m}' for m in range(self.n_enc)] if self.mulmo else ['enc']
        self.enc_names = enc_names
        self.filters = []
        f = self.F
        for _ in range(self.n):                               # components.py:204-220
            self.filters.append(f)
            f = int(self.rate * f)
        for en in enc_names:
            cin = 1 if self.mulmo else cin_total
            for i, f in enumerate(self.filters):             # Downsample, components.py:46-61
                for k in range(self.n_conv):
                    self._add_conv(rng, f'{en}/d{i}/conv{k}', self.k, self.k, cin, f)
                    if self.bn:
                        self._add_bn(f'{en}/d{i}/bn{k}', f)
                    cin = f
                if self.bn:
                    self._add_bn(f'{en}/d{i}/pool_bn', f)
        cin = self.filters[-1] * self.n_enc                   # unet.py:172-175
        for j, rc in enumerate(reversed(self.filters)):       # Decoder.build, components.py:293-306
            self._add_tconv(rng, f'dec/u{j}/tconv', self.rate, cin, rc)
            if self.bn:
                self._add_bn(f'dec/u{j}/tconv_bn', rc)
            c = 2 * rc
            for k in range(self.n_conv):
                self._add_conv(rng, f'dec/u{j}/conv{k}', self.k, self.k, c, rc)
                if self.bn:
                    self._add_bn(f'dec/u{j}/bn{k}', rc)
                c = rc
            cin = rc
        self._add_conv(rng, 'head', 1, 1, cin, 1)             # unet.py:241-244
        return self

    def _encoder(self, ctx, en, x):
        w = ctx['w']
        res = []
        for i in range(self.n):                               # Encoder.call, components.py:235-247
            for k in range(self.n_conv):                      # Downsample.call :77-81
                x = ops.conv2d(x, w[f'{en}/d{i}/conv{k}/kernel'], w[f'{en}/d{i}/conv{k}/bias'], self.padding)
                x = ops.activation(x, self.act)
                if self.bn:
                    x = self._bn(ctx, f'{en}/d{i}/bn{k}', x)
            res.append(x)
            ctx['tensors'][f'{en}/d{i}/res'] = x
            x, idx = ops.maxpool(x, self.rate, return_indices=True)
            ctx['pool_idx'].append(idx)
            if self.bn:
                x = self._bn(ctx, f'{en}/d{i}/pool_bn', x)
        return res, x

    def _graph(self, ctx, x):
        w = ctx['w']
        if self.mulmo:                                        # MulmoUNet.call, unet.py:180-191
            outs = [self._encoder(ctx, en, x[..., m:m + 1]) for m, en in enumerate(self.enc_names)]
            res = outs[self.reference_index][0]
            x = torch.cat([o[1] for o in outs], dim=-1)
        else:                                                 # UNet.call, unet.py:84-88
            res, x = self._encoder(ctx, 'enc', x)
        ctx['tensors']['bottleneck'] = x
        for j, ref in enumerate(reversed(res)):               # Decoder.call, components.py:314-320
            t = ops.conv2d_transpose(x, w[f'dec/u{j}/tconv/kernel'], w[f'dec/u{j}/tconv/bias'], self.rate)
            if self.bn:
                t = self._bn(ctx, f'dec/u{j}/tconv_bn', t)
            gh = (ref.shape[1] - t.shape[1]) // 2             # Upsample.call, components.py:160-164
            gw = (ref.shape[2] - t.shape[2]) // 2
            cropped = ref[:, gh:gh + t.shape[1], gw:gw + t.shape[2], :]
            x = torch.cat([t, cropped], dim=-1)
            for k in range(self.n_conv):
                x = ops.conv2d(x, w[f'dec/u{j}/conv{k}/kernel'], w[f'dec/u{j}/conv{k}/bias'], self.padding)
                x = ops.activation(x, self.act)
                if self.bn:
                    x = self._bn(ctx, f'dec/u{j}/bn{k}', x)
            ctx['tensors'][f'dec/u{j}/out'] = x
        # last_conv: kernel_size=1, sigmoid (applied by the caller on the cached logits)
        return ops.conv2d(x, w['head/kernel'], w['head/bias'], self.padding)


class RefMultiResUnet(_Base):
    """``MultiResUnet(height, width, n_channels)`` (multiresunet.py:167-223)."""

    def __init__(self, height=None, width=None, n_channels=5):
        super().__init__()
        self.n_channels = n_channels

    def build(self, input_shape=None, seed=0, dtype=torch.float32):
        self.dtype = dtype
        self._rng = np.random.default_rng(seed)
        self._building = True
        self._counters = dict(conv=0, bn=0, tconv=0)
        ctx = dict(w=self.weights, training=False, new_moving=OrderedDict(), pool_idx=[], tensors=OrderedDict())
        self._graph(ctx, torch.zeros(1, 16, 16, self.n_channels, dtype=dtype))
        self._building = False
        return self

    # layer factories: create on first (build) pass, look up by creation order afterwards
    def _next(self, kind):
        n = self._counters[kind]
        self._counters[kind] = n + 1
        return f'{kind}{n}'

    def _conv2d_bn(self, ctx, x, filters, k, activation='relu'):
        """multiresunet.py:31-60: Conv2D(use_bias=False) -> BN(scale=False) -> activation."""
        cname, bname = self._next('conv'), self._next('bn')
        if self._building:
            self._add_conv(self._rng, cname, k, k, x.shape[-1], filters, bias=False)
            self._add_bn(bname, filters, scale=False)
        x = ops.conv2d(x, ctx['w'][f'{cname}/kernel'], None, 'same')
        x = self._bn(ctx, bname, x)
        if activation == 'sigmoid':
            return x  # caller applies the sigmoid on the logits
        return ops.activation(x, activation)

    def _full_bn(self, ctx, x):
        bname = self._next('bn')
        if self._building:
            self._add_bn(bname, x.shape[-1], scale=True)
        return self._bn(ctx, bname, x)

    def _tconv(self, ctx, x, filters):
        name = self._next('tconv')
        if self._building:
            self._add_tconv(self._rng, name, 2, x.shape[-1], filters)
        return ops.conv2d_transpose(x, ctx['w'][f'{name}/kernel'], ctx['w'][f'{name}/bias'], 2)

    def _mres_block(self, ctx, U, inp, alpha=1.67):
        """multiresunet.py:89-126."""
        W = alpha * U
        f1, f2, f3 = int(W * 0.167), int(W * 0.333), int(W * 0.5)
        shortcut = self._conv2d_bn(ctx, inp, f1 + f2 + f3, 1, activation=None)
        c3 = self._conv2d_bn(ctx, inp, f1, 3)
        c5 = self._conv2d_bn(ctx, c3, f2, 3)
        c7 = self._conv2d_bn(ctx, c5, f3, 3)
        out = torch.cat([c3, c5, c7], dim=-1)
        out = self._full_bn(ctx, out)
        out = torch.relu(shortcut + out)
        return self._full_bn(ctx, out)

    def _res_path(self, ctx, filters, length, inp):
        """multiresunet.py:129-164."""
        out = inp
        for _ in range(length):
            shortcut = self._conv2d_bn(ctx, out, filters, 1, activation=None)
            o = self._conv2d_bn(ctx, out, filters, 3)
            out = torch.relu(shortcut + o)
            out = self._full_bn(ctx, out)
        return out

    def _graph(self, ctx, x):
        self._counters = dict(conv=0, bn=0, tconv=0)
        skips = []
        for lvl, length in enumerate((4, 3, 2, 1)):           # multiresunet.py:182-196
            b = self._mres_block(ctx, 32 * 2 ** lvl, x)
            x = ops.maxpool(b, 2)
            skips.append(self._res_path(ctx, 32 * 2 ** lvl, length, b))
        x = self._mres_block(ctx, 32 * 16, x)                 # :198
        for lvl in (3, 2, 1, 0):                              # :200-217
            up = torch.cat([self._tconv(ctx, x, 32 * 2 ** lvl), skips[lvl]], dim=-1)
            x = self._mres_block(ctx, 32 * 2 ** lvl, up)
        return self._conv2d_bn(ctx, x, 1, 1, activation='sigmoid')  # :219


def build_model(name, model_options, input_shape, seed=0, dtype=torch.float32):
    """engine.py:267-268 ``getattr(tf_models, model_name)(**model_options)`` + ``build``."""
    if name == 'UNetAnnotator':
        m = RefUNetAnnotator(**model_options)
    elif name == 'MulmoUNetAnnotator':
        m = RefUNetAnnotator(**model_options, mulmo=True)
    elif name == 'MultiResUnet':
        m = RefMultiResUnet(**model_options)
    else:
        raise ValueError(name)
    return m.build(input_shape, seed=seed, dtype=dtype)
