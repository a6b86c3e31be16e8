"""UNet / MulmoUNet / UNetAnnotator / MulmoUNetAnnotator with the reference's
names and constructor arguments (``annotator/models/tf_models/unet.py``).

``engine.py:267-268`` builds models with ``getattr(tf_models, name)(**model_options)``;
the YAML ``model_options`` (configs/unet.yaml, unet_big.yaml, mulmo_unet.yaml) map
1:1 onto these constructors.
"""
from __future__ import annotations

from ... import runtime as R
from ...keras_like import Model
from . import components
from .components import Layer


class UNet(Layer):
    '''U-Net (unet.py:19-88): decoder(encoder(x))'''

    def __init__(self, filters_first, n_downsample, rate, kernel_size, conv_stride, bn=False, trainable=True,
                 padding='valid', activation='relu', kernel_regularizer=None, **kargs):
        super().__init__(**kargs)
        self.configs = dict(filters_first=filters_first, n_downsample=n_downsample, rate=rate,
                            kernel_size=kernel_size, conv_stride=conv_stride, bn=bn, trainable=trainable,
                            padding=padding, activation=activation, kernel_regularizer=kernel_regularizer)
        common = dict(rate=rate, kernel_size=kernel_size, conv_stride=conv_stride, bn=bn, padding=padding,
                      activation=activation, trainable=trainable, kernel_regularizer=kernel_regularizer)
        self.encoder = components.Encoder(filters_first=filters_first, n_downsample=n_downsample, name='enc', **common)
        self.decoder = components.Decoder(name='dec', **common)

    def get_config(self):
        return dict(self.configs)

    def bind(self, ctx):
        super().bind(ctx)
        self.encoder.bind(ctx)
        self.decoder.bind(ctx)
        return self

    def build(self, input_shape):
        self.encoder_output_shape, self.ref_shapes = self.encoder.build(input_shape)
        decoder_out = self.decoder.build(self.encoder_output_shape, self.ref_shapes)
        self.built = True
        return decoder_out

    def emit(self, plan, x):
        res_list, down = self.encoder.emit(plan, x)
        return self.decoder.emit(plan, down, res_list)


class MulmoUNet(Layer):
    '''MulmoU-Net (unet.py:91-191): one encoder per input channel, bottlenecks concatenated,
    skip connections from the ``reference_index`` encoder only.'''

    def __init__(self, filters_first, n_downsample, rate, kernel_size, conv_stride, bn=False, trainable=True,
                 padding='valid', activation='relu', kernel_regularizer=None, reference_index=0, **kargs):
        super().__init__(**kargs)
        self.configs = dict(filters_first=filters_first, n_downsample=n_downsample, rate=rate,
                            kernel_size=kernel_size, conv_stride=conv_stride, bn=bn, trainable=trainable,
                            padding=padding, activation=activation, kernel_regularizer=kernel_regularizer,
                            reference_index=reference_index)
        self.decoder = components.Decoder(rate=rate, kernel_size=kernel_size, conv_stride=conv_stride, bn=bn,
                                          padding=padding, activation=activation, trainable=trainable,
                                          kernel_regularizer=kernel_regularizer, name='dec')
        self.reference_index = reference_index
        self.encoders = []

    def get_config(self):
        return dict(self.configs)

    def build(self, input_shape):
        c = self.configs
        self.channel_len = input_shape[-1]
        self.encoders = [
            components.Encoder(filters_first=c['filters_first'], n_downsample=c['n_downsample'], rate=c['rate'],
                               kernel_size=c['kernel_size'], conv_stride=c['conv_stride'], bn=c['bn'],
                               padding=c['padding'], activation=c['activation'], trainable=c['trainable'],
                               kernel_regularizer=c['kernel_regularizer'], name=f'enc/{m}').bind(self._ctx)
            for m in range(self.channel_len)]
        self.decoder.bind(self._ctx)
        outs = [enc.build((*input_shape[:-1], 1)) for enc in self.encoders]          # unet.py:166-168
        self.encoder_output_shape_list = [o[0] for o in outs]
        self.ref_shapes_list = [o[1] for o in outs]
        last_dim = sum(s[-1] for s in self.encoder_output_shape_list)
        self.ref_shapes = self.ref_shapes_list[self.reference_index]
        self.encoder_output_shape = (*self.encoder_output_shape_list[0][:3], last_dim)
        decoder_out = self.decoder.build(self.encoder_output_shape, self.ref_shapes)
        self.built = True
        return decoder_out

    def emit(self, plan, x):
        fb = self.encoder_output_shape_list[0][-1]
        n = len(self.ref_shapes)
        x0 = x[0] if isinstance(x, list) else x
        bott = plan.new_buf(x0.h >> n, x0.w >> n, fb * self.channel_len, 'bottleneck')   # tf.concat, unet.py:187
        res_ref = None
        for m, enc in enumerate(self.encoders):
            if isinstance(x, list):
                xin = x[m]
            else:
                xin = R.TRef(x.buf, x.coff + m, 1)                                        # inputs[..., m:m+1]
                xin.needs_grad = plan.want_input_grad
            op0 = len(plan.ops)
            res_list, _ = enc.emit(plan, xin, out_dst=R.TRef(bott, m * fb, fb))
            plan.branches.append((op0, len(plan.ops)))          # the per-modality encoders are independent of each other
            if m == self.reference_index:
                res_ref = res_list
        return self.decoder.emit(plan, R.TRef(bott), res_ref)


class UNetAnnotator(Model):
    """``UNetAnnotator`` (unet.py:194-282): UNet + Conv2D(1, k=1, sigmoid) head.  The head and the
    sigmoid run fused with the loss (``dnnca_head_bce_fwd_bwd``) or as ``dnnca_head_fwd``."""

    def __init__(self, n_filters_first, n_downsample, rate, kernel_size, conv_stride, bn=False, padding='valid',
                 activation='relu', kernel_regularizer=None, dtype=None, seed=0, **kargs):
        super().__init__(dtype=dtype, seed=seed)
        self.configs = dict(n_filters_first=n_filters_first, n_downsample=n_downsample, rate=rate,
                            kernel_size=kernel_size, conv_stride=conv_stride, bn=bn, padding=padding,
                            activation=activation, kernel_regularizer=kernel_regularizer, **kargs)
        self.kargs = kargs
        self.padding = padding
        self.unet = self.construct_internal_model().bind(self._ctx)

    def construct_internal_model(self):
        c = self.configs
        return UNet(filters_first=c['n_filters_first'], n_downsample=c['n_downsample'], rate=c['rate'],
                    kernel_size=c['kernel_size'], conv_stride=c['conv_stride'], bn=c['bn'], padding=c['padding'],
                    activation=components.solve_activation(c['activation']),
                    kernel_regularizer=c['kernel_regularizer'], **self.kargs)

    def get_config(self):
        return dict(self.configs)

    @classmethod
    def from_config(cls, config):
        return cls(**config)

    def _build_variables(self, input_shape):
        unet_out = self.unet.build(input_shape)
        l2 = components.solve_regularizer(self.configs['kernel_regularizer'])
        self.unet._add_conv('head', 1, 1, unet_out[-1], 1, l2)       # last_conv, unet.py:241-244

    def _emit(self, plan):
        x = plan.input
        if plan.dtype != x.buf.dtype:                                 # fp32 input -> bf16 activations
            wide = self.configs['n_filters_first'] % 16 == 0          # first conv runs on the tensor cores
            if wide and isinstance(self.unet, MulmoUNet):
                # one dense single-channel buffer per modality (unet.py:182-185 slices inputs[..., m:m+1]): the
                # 1 -> F first convs then run on the row-Toeplitz tcgen05 kernels (conv_row_umma.cu)
                xs = []
                for m in range(x.c):
                    xm = R.TRef(plan.new_buf(x.h, x.w, 1, f'input_cast{m}', zero=True), 0, 1)
                    xm.needs_grad = plan.want_input_grad
                    src = R.TRef(x.buf, m, 1)
                    src.needs_grad = plan.want_input_grad
                    plan.add(R.ConvertOp(plan, src, xm))
                    xs.append(xm)
                x = xs
            else:
                cpad = (x.c + 7) // 8 * 8 if wide else x.c            # 16-byte aligned pixels for the TMA
                xb = R.TRef(plan.new_buf(x.h, x.w, cpad, 'input_cast', zero=True), 0, x.c)
                xb.needs_grad = plan.want_input_grad
                plan.add(R.ConvertOp(plan, x, xb))
                x = xb
        plan.features = self.unet.emit(plan, x)
        plan.head = ('head/kernel', 'head/bias')


class MulmoUNetAnnotator(UNetAnnotator):
    '''Annotator model based on MulmoUNet (unet.py:285-300)'''

    def construct_internal_model(self):
        c = self.configs
        return MulmoUNet(filters_first=c['n_filters_first'], n_downsample=c['n_downsample'], rate=c['rate'],
                         kernel_size=c['kernel_size'], conv_stride=c['conv_stride'], bn=c['bn'],
                         padding=c['padding'], activation=components.solve_activation(c['activation']),
                         kernel_regularizer=c['kernel_regularizer'], **self.kargs)
