"""Building blocks with the reference's names and constructor arguments
(``annotator/models/tf_models/components.py``), lowered to libdnnca ops.

Each class keeps the reference's two-phase protocol: ``build(input_shape)``
creates the variables and propagates shapes (components.py:69-75,142-151,
226-233,292-312); ``emit(plan, ...)`` is the counterpart of ``call`` and
records the ops of one forward pass into a static ``runtime.Plan``.
"""
from __future__ import annotations

import numpy as np

from ... import native as N
from ... import runtime as R


def solve_activation(identifier):
    """components.py:323-335.  Returns ``(act_code, alpha)``.

    Accepts the strings/dicts the YAML surface uses: ``'relu'``, ``None``/``'linear'``
    and ``{class_name: LeakyReLU, config: {alpha: a}}`` (configs/additionals/leakyReLU.yaml)."""
    if isinstance(identifier, tuple) and len(identifier) == 2:
        return identifier
    if identifier is None or identifier == 'linear':
        return (N.ACT_NONE, 0.0)
    if isinstance(identifier, str):
        if identifier == 'relu':
            return (N.ACT_RELU, 0.0)
        raise ValueError(f'Failed to resolve activation: {identifier}')
    if isinstance(identifier, dict):
        cls = identifier.get('class_name')
        cfg = identifier.get('config', {}) or {}
        if cls == 'LeakyReLU':
            return (N.ACT_LEAKY, float(cfg.get('alpha', 0.3)))
        if cls == 'ReLU':
            return (N.ACT_RELU, 0.0)
    raise ValueError(f'Failed to resolve activation: {identifier}')


def solve_regularizer(identifier):
    """configs/additionals/kernel_regularizer.yaml:1-4 -> L2 coefficient (0 = none)."""
    if identifier is None:
        return 0.0
    if isinstance(identifier, dict) and identifier.get('class_name') in ('L2', 'l2'):
        return float((identifier.get('config') or {}).get('l2', 0.01))
    if isinstance(identifier, str) and identifier.lower() == 'l2':
        return 0.01
    raise ValueError(f'unsupported kernel_regularizer: {identifier}')


def glorot_uniform(rng, shape):
    """keras default kernel initialiser: limit = sqrt(6 / (fan_in + fan_out))."""
    rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    limit = np.sqrt(6.0 / ((shape[-2] + shape[-1]) * rf))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


class Layer:
    """Minimal stand-in for ``keras.layers.Layer``: a name scope + the shared ParamStore/rng."""

    def __init__(self, name=None, **kargs):
        if kargs:       # keras.layers.Layer would reject them too; silently dropping a config key hides typos
            raise TypeError(f'{type(self).__name__}: unexpected keyword arguments {sorted(kargs)}')
        self.name = name or type(self).__name__.lower()
        self.built = False
        self._ctx = None

    def bind(self, ctx):
        """ctx carries the ParamStore and the initialiser rng shared by the whole model."""
        self._ctx = ctx
        return self

    # variable factories ------------------------------------------------------
    def _add_conv(self, prefix, kh, kw, cin, cout, l2, bias=True):
        ps, rng = self._ctx['params'], self._ctx['rng']
        ps.add(f'{prefix}/kernel', glorot_uniform(rng, (kh, kw, cin, cout)), l2=l2)
        if bias:
            ps.add(f'{prefix}/bias', np.zeros(cout, np.float32))

    def _add_tconv(self, prefix, k, cin, cout, l2):
        ps, rng = self._ctx['params'], self._ctx['rng']
        ps.add(f'{prefix}/kernel', glorot_uniform(rng, (k, k, cout, cin)), l2=l2)
        ps.add(f'{prefix}/bias', np.zeros(cout, np.float32))

    def _add_bn(self, prefix, c, scale=True):
        ps = self._ctx['params']
        if scale:
            ps.add(f'{prefix}/gamma', np.ones(c, np.float32))
        ps.add(f'{prefix}/beta', np.zeros(c, np.float32))
        ps.add(f'{prefix}/moving_mean', np.zeros(c, np.float32), trainable=False)
        ps.add(f'{prefix}/moving_var', np.ones(c, np.float32), trainable=False)


def _check_supported(rate, kernel_size, conv_stride, padding, trainable=True):
    if not trainable:
        raise NotImplementedError(
            'trainable=False (frozen Conv/BatchNormalization layers, components.py:47-61) is not implemented: the fused '
            'optimizer updates every variable of the flat parameter buffer')
    if padding != 'same':
        raise NotImplementedError(
            "padding='valid' is accepted by the reference code (components.py:161-163) but used by none of its "
            "configs; the B200 kernels implement padding='same' only")
    if rate != 2 or conv_stride != 1 or kernel_size not in (1, 3):
        raise NotImplementedError(
            f'unsupported geometry rate={rate} kernel_size={kernel_size} conv_stride={conv_stride}: the kernels cover '
            'the reference configs (rate 2, 3x3 stride-1 convs)')


def conv_act_bn(layer, plan, x, dst, prefix, bnprefix, ksize, act, bn, bias=True, x2=None):
    """Conv2D(+bias, activation) [-> BatchNormalization]   (components.py:46-61, 122-134).

    Without BN the conv writes straight into ``dst``; with BN the conv writes its
    (post-activation) output to a scratch tensor while accumulating the batch statistics in its
    epilogue, and the BN apply pass writes ``dst``."""
    kernel, b = f'{prefix}/kernel', (f'{prefix}/bias' if bias else None)
    if not bn:
        plan.add(R.ConvOp(plan, x, dst, kernel, b, ksize, act, x2=x2))
        return dst
    a = R.TRef(plan.new_buf(x.h, x.w, dst.c, prefix + ':a'))
    st = R.BNStats(plan, dst.c)
    plan.add(R.ConvOp(plan, x, a, kernel, b, ksize, act, stats=st, x2=x2))
    plan.add(R.BNOp(plan, a, dst, bnprefix, st))
    return dst


class Downsample(Layer):
    '''downsampling block (components.py:16-81)'''

    def __init__(self, filters, rate, kernel_size, conv_stride, bn, n_conv=2, trainable=True, padding='valid',
                 activation='relu', kernel_regularizer=None, **kargs):
        super().__init__(**kargs)
        _check_supported(rate, kernel_size, conv_stride, padding, trainable)
        self.configs = dict(filters=filters, rate=rate, kernel_size=kernel_size, conv_stride=conv_stride, bn=bn,
                            n_conv=n_conv, trainable=trainable, padding=padding, activation=activation,
                            kernel_regularizer=kernel_regularizer)
        self.filters, self.rate, self.kernel_size, self.bn, self.n_conv = filters, rate, kernel_size, bn, n_conv
        self.act = solve_activation(activation)
        self.l2 = solve_regularizer(kernel_regularizer)

    def get_config(self):
        return dict(self.configs)

    def build(self, input_shape):
        cin = input_shape[-1]
        for k in range(self.n_conv):
            self._add_conv(f'{self.name}/conv{k}', self.kernel_size, self.kernel_size, cin, self.filters, self.l2)
            if self.bn:
                self._add_bn(f'{self.name}/bn{k}', self.filters)
            cin = self.filters
        if self.bn:
            self._add_bn(f'{self.name}/pool_bn', self.filters)
        conv_output_shape = (*input_shape[:3], self.filters)
        pool_output_shape = (input_shape[0], input_shape[1] // self.rate, input_shape[2] // self.rate, self.filters)
        self.built = True
        return conv_output_shape, pool_output_shape

    def emit(self, plan, x, res_dst=None, half_dst=None):
        """-> (conv, half) like ``call`` (components.py:77-81).  ``res_dst`` / ``half_dst`` let the
        caller place an output inside a concat buffer (MulmoUNet's bottleneck, unet.py:187)."""
        f = self.filters
        if x.h % 2 or x.w % 2:
            raise ValueError(f'{self.name}: spatial size {x.h}x{x.w} is not divisible by the pool rate')
        for k in range(self.n_conv):
            last = k == self.n_conv - 1
            dst = res_dst if (last and res_dst is not None) else R.TRef(plan.new_buf(x.h, x.w, f, f'{self.name}/conv{k}'))
            x = conv_act_bn(self, plan, x, dst, f'{self.name}/conv{k}', f'{self.name}/bn{k}', self.kernel_size,
                            self.act, self.bn)
        conv = x
        half = half_dst if (half_dst is not None and not self.bn) else R.TRef(
            plan.new_buf(x.h // 2, x.w // 2, f, f'{self.name}/pool'))
        if not self.bn:
            plan.add(R.PoolOp(plan, conv, half))
            return conv, half
        st = R.BNStats(plan, f)
        plan.add(R.PoolOp(plan, conv, half, stats=st))
        out = half_dst if half_dst is not None else R.TRef(plan.new_buf(half.h, half.w, f, f'{self.name}/pool_bn'))
        plan.add(R.BNOp(plan, half, out, f'{self.name}/pool_bn', st))
        return conv, out


class Upsample(Layer):
    """upsampling block (components.py:84-166)"""

    def __init__(self, filters, rate, kernel_size, conv_stride, bn, trainable, n_conv=2, padding='valid',
                 activation='relu', kernel_regularizer=None, **kargs):
        super().__init__(**kargs)
        _check_supported(rate, kernel_size, conv_stride, padding, trainable)
        self.configs = dict(filters=filters, rate=rate, kernel_size=kernel_size, conv_stride=conv_stride, bn=bn,
                            trainable=trainable, n_conv=n_conv, padding=padding, activation=activation,
                            kernel_regularizer=kernel_regularizer)
        self.filters, self.rate, self.kernel_size, self.bn, self.n_conv = filters, rate, kernel_size, bn, n_conv
        self.act = solve_activation(activation)
        self.l2 = solve_regularizer(kernel_regularizer)

    def get_config(self):
        return dict(self.configs)

    def build(self, input_shape, ref_shape):
        self._add_tconv(f'{self.name}/tconv', self.rate, input_shape[-1], self.filters, self.l2)
        if self.bn:
            self._add_bn(f'{self.name}/tconv_bn', self.filters)
        cin = self.filters + ref_shape[-1]
        for k in range(self.n_conv):
            self._add_conv(f'{self.name}/conv{k}', self.kernel_size, self.kernel_size, cin, self.filters, self.l2)
            if self.bn:
                self._add_bn(f'{self.name}/bn{k}', self.filters)
            cin = self.filters
        self.built = True

    def compute_output_shape(self, input_shape, ref_shape):
        return [input_shape[0], input_shape[1] * self.rate, input_shape[2] * self.rate, self.filters]

    def emit(self, plan, x, reference):
        """``tf.concat([tconv0, cropped], -1)`` (components.py:164: tconv first, skip second) is
        virtual: the first conv reads the transposed-conv output and the skip tensor as two inputs.
        The centre crop (components.py:161-163) is the identity under padding='same'."""
        f = self.filters
        assert (reference.h, reference.w) == (x.h * self.rate, x.w * self.rate)
        t = R.TRef(plan.new_buf(reference.h, reference.w, f, f'{self.name}/tconv'))
        if not self.bn:
            plan.add(R.TConvOp(plan, x, t, f'{self.name}/tconv/kernel', f'{self.name}/tconv/bias'))
        else:
            st = R.BNStats(plan, f)
            plan.add(R.TConvOp(plan, x, t, f'{self.name}/tconv/kernel', f'{self.name}/tconv/bias', stats=st))
            tb = R.TRef(plan.new_buf(reference.h, reference.w, f, f'{self.name}/tconv_bn'))
            plan.add(R.BNOp(plan, t, tb, f'{self.name}/tconv_bn', st))
            t = tb
        reference.skip_consumed = True       # its gradient arrives from this conv AND from its max-pool
        x, x2 = t, reference
        for k in range(self.n_conv):
            dst = R.TRef(plan.new_buf(x.h, x.w, f, f'{self.name}/conv{k}'))
            x = conv_act_bn(self, plan, x, dst, f'{self.name}/conv{k}', f'{self.name}/bn{k}', self.kernel_size,
                            self.act, self.bn, x2=x2)
            x2 = None
        return x


class Encoder(Layer):
    """encoder block (components.py:169-247)"""

    def __init__(self, filters_first, n_downsample, rate, kernel_size, conv_stride, bn, trainable, n_conv=2,
                 padding='valid', activation='relu', kernel_regularizer=None, **kargs):
        super().__init__(**kargs)
        self.configs = dict(filters_first=filters_first, n_downsample=n_downsample, rate=rate,
                            kernel_size=kernel_size, conv_stride=conv_stride, bn=bn, trainable=trainable,
                            n_conv=n_conv, padding=padding, activation=activation,
                            kernel_regularizer=kernel_regularizer)
        self.downsamples = []
        next_filters = filters_first
        for i in range(n_downsample):
            self.downsamples.append(Downsample(
                filters=next_filters, rate=rate, kernel_size=kernel_size, conv_stride=conv_stride, n_conv=n_conv,
                bn=bn, padding=padding, trainable=trainable, activation=activation,
                kernel_regularizer=kernel_regularizer, name=f'{self.name}/d{i}'))
            next_filters = int(rate * next_filters)     # components.py:220

    def get_config(self):
        return dict(self.configs)

    def bind(self, ctx):
        super().bind(ctx)
        for d in self.downsamples:
            d.bind(ctx)
        return self

    def build(self, input_shape):
        ref_shapes = []
        output_shape = input_shape
        for downsample in self.downsamples:
            ref_shape, output_shape = downsample.build(output_shape)
            ref_shapes.append(ref_shape)
        self.built = True
        return output_shape, ref_shapes

    def emit(self, plan, x, res_dsts=None, out_dst=None):
        """-> (res_list, downsampled) like ``call`` (components.py:235-247)."""
        res_list = []
        for i, d in enumerate(self.downsamples):
            last = i == len(self.downsamples) - 1
            res, x = d.emit(plan, x, res_dst=res_dsts[i] if res_dsts else None, half_dst=out_dst if last else None)
            res_list.append(res)
        return res_list, x


class Decoder(Layer):
    """decoder block (components.py:250-320)"""

    def __init__(self, rate, kernel_size, conv_stride, bn, trainable, padding='valid', activation='relu',
                 kernel_regularizer=None, **kargs):
        super().__init__(**kargs)
        self.configs = dict(rate=rate, kernel_size=kernel_size, conv_stride=conv_stride, bn=bn, trainable=trainable,
                            padding=padding, activation=activation, kernel_regularizer=kernel_regularizer)
        self.upsamples = []

    def get_config(self):
        return dict(self.configs)

    def build(self, inputs_shape, ref_shapes):
        c = self.configs
        for j, ref_shape in enumerate(reversed(ref_shapes)):     # components.py:293-306
            up = Upsample(filters=ref_shape[-1], rate=c['rate'], kernel_size=c['kernel_size'],
                          conv_stride=c['conv_stride'], bn=c['bn'], trainable=c['trainable'], padding=c['padding'],
                          activation=c['activation'], kernel_regularizer=c['kernel_regularizer'],
                          name=f'{self.name}/u{j}').bind(self._ctx)
            self.upsamples.append(up)
            up.build(inputs_shape, ref_shape)
            inputs_shape = up.compute_output_shape(inputs_shape, ref_shape)
        self.built = True
        return inputs_shape

    def emit(self, plan, x, res_list):
        assert len(res_list) == len(self.upsamples), \
            f'#References {len(res_list)} != #upsamples {len(self.upsamples)}'
        for reference, up in zip(reversed(res_list), self.upsamples):
            x = up.emit(plan, x, reference)
        return x
