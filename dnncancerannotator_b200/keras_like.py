"""The slice of the ``tf.keras.Model`` surface that the reference's engine and
callbacks consume (``engine.py:93,126-135,198-203,222,286``;
``callbacks.py:290-303``), implemented over a static ``runtime.Plan``.

``Model.train_step`` is the hot path: H2D copy of the batch into static buffers,
then ONE CUDA-graph replay of {zero grads/stats, forward, fused head+loss,
backward, (gradient all-reduce), fused Adam}.
"""
from __future__ import annotations

import os
import re
from collections import OrderedDict

import numpy as np
import torch

from . import native as N
from . import runtime as R
from .utils import losses as L


class History:
    """What ``Model.fit`` returns (``dump.py:68-73`` reads .epoch/.history/.params)."""

    def __init__(self, model, params):
        self.model, self.params = model, params
        self.epoch, self.history = [], {}

    def record(self, epoch, logs):
        self.epoch.append(epoch)
        for k, v in logs.items():
            self.history.setdefault(k, []).append(float(v))


def _to_device_unit(a, device):
    """The ONE input-normalisation door of every entry point (``__call__``, ``predict``, ``evaluate``,
    ``forward_backward``, ``train_step``): float inputs are the reference's contract (float32 in [0,1],
    data.py:193-206) and are only cast; raw uint8 slices / labels are divided by 255 ON THE DEVICE
    (``dnnca_u8_to_unit``, data.py:206), so a model trained on uint8 batches sees the same values when it is
    evaluated or asked to predict on them."""
    t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype == torch.uint8:
        src = t.to(device, non_blocking=True).contiguous()
        out = torch.empty(src.shape, dtype=torch.float32, device=device)
        N.call('dnnca_u8_to_unit', N.stream_ptr(), N.ptr(src), src.numel(), N.ptr(out), N.F32)
        return out
    if t.dtype not in (torch.float32, torch.float64, torch.float16, torch.bfloat16):
        raise TypeError(f'inputs must be float in [0,1] or raw uint8 slices, got {t.dtype}')
    return t.to(device=device, dtype=torch.float32, non_blocking=True)


class LoadStatus:
    """What ``load_weights`` returns: the slice of TF's ``CheckpointLoadStatus`` that ``engine.py:75`` calls."""

    def __init__(self, model_names, file_names):
        self.missing = sorted(set(model_names) - set(file_names))       # model variables the file does not hold
        self.unused = sorted(set(file_names) - set(model_names))        # file entries no model variable claimed

    def assert_existing_objects_matched(self):
        if self.missing:
            raise AssertionError(f'checkpoint holds no value for {len(self.missing)} model variables: {self.missing[:5]}...')
        return self

    def assert_consumed(self):
        self.assert_existing_objects_matched()
        if self.unused:
            raise AssertionError(f'{len(self.unused)} checkpoint entries were not used: {self.unused[:5]}...')
        return self

    def expect_partial(self):
        return self


def _is_packed(y):
    from .data_tail import PackedLabels
    return isinstance(y, PackedLabels)


def _label_tensor(y):
    """labels as a torch tensor: float32 / uint8 ``[B,H,W]``, or the bits of ``data_tail.PackedLabels``"""
    if _is_packed(y):
        return y.bits
    return y if torch.is_tensor(y) else torch.from_numpy(np.ascontiguousarray(y))


def load_model(path, dtype=None):
    """Rebuilds what ``Model.save(path)`` wrote (the counterpart of ``tf.keras.models.load_model`` for the directory of
    engine.py:226): class + constructor config from ``config.json``, compiled with the saved loss / optimizer settings,
    variables and optimizer slots from ``weights.npz``.  ``dtype`` overrides the saved compute dtype."""
    import json
    from .models import tf_models
    with open(os.path.join(path, 'config.json')) as f:
        cfg = json.load(f)
    cls = getattr(tf_models, cfg['class_name'], None)
    if cls is None or not (isinstance(cls, type) and issubclass(cls, Model)):
        raise ValueError(f"unknown model class {cfg['class_name']!r} in {path}/config.json")
    model = cls(**cfg.get('config', {}), dtype=dtype or cfg.get('compute_dtype'))
    shape = cfg.get('input_shape')
    if not shape:
        raise ValueError(f'{path}/config.json holds no input shape: the model was saved before it was built')
    model.build(tuple(shape))
    model.compile(optimizer=cfg.get('optimizer') or 'adam', loss=cfg.get('loss'))
    if torch.cuda.is_available():
        model.params.materialize(model.device)       # so that the Adam slots and the step counter are restored as well
    model.load_weights(path).assert_existing_objects_matched()
    return model


class Model:
    """Base of ``UNetAnnotator`` / ``MulmoUNetAnnotator`` / ``MultiResUnet``.

    ``compute_dtype``: 'bf16' (default: bf16 activations, fp32 accumulate, fp32 master weights) or
    'fp32' (bit-exact-mask mode).  Chosen with the ``dtype=`` constructor keyword or the
    ``DNNCA_DTYPE`` environment variable; it is not part of the reference's config surface.
    """

    def __init__(self, dtype=None, seed=0, device=None, **kargs):
        if kargs:
            raise TypeError(f'unexpected keyword arguments {sorted(kargs)}')
        dtype = dtype or os.environ.get('DNNCA_DTYPE', 'bf16')
        self.compute_dtype = {'bf16': torch.bfloat16, 'bfloat16': torch.bfloat16, 'fp32': torch.float32,
                              'float32': torch.float32}[dtype]
        self.params = R.ParamStore()
        self._ctx = dict(params=self.params, rng=np.random.default_rng(seed))
        self._device = device
        self._plans = {}
        self.built = False
        self.input_shape = None
        self.loss = None
        self.optimizer = None
        self.metrics = []
        self._metric_set = None
        self.stop_training = False
        self.use_cuda_graph = os.environ.get('DNNCA_NO_GRAPH', '0') != '1'
        self.trainable_model = True     # MultiResUnet (inference-only in this build) sets False
        # data parallel (engine.py:260-263 MirroredStrategy): set by enable_data_parallel()
        self._dp = None
        self._p2p = None

    # ---- to be provided by subclasses ------------------------------------------
    def _build_variables(self, input_shape):
        raise NotImplementedError

    def _emit(self, plan):
        raise NotImplementedError

    # ---- keras surface ---------------------------------------------------------
    @property
    def device(self):
        if self._device is None:
            if not torch.cuda.is_available():
                raise N.DnncaError('no CUDA device: dnncancerannotator_b200 has no CPU fallback')
            self._device = torch.device('cuda', torch.cuda.current_device())
        return self._device

    def build(self, input_shape):
        """engine.py:93 ``model.build((None, H, W, C))``: creates the variables."""
        if self.built:
            return
        assert len(input_shape) == 4
        self.input_shape = tuple(input_shape)
        self._build_variables(self.input_shape)
        self.built = True

    def compile(self, optimizer='adam', loss=None, metrics=None, **kargs):
        """engine.py:286.  ``loss``: a ``WeightedCrossentropy`` instance or its keras-style
        dict/str identifier (deploy_options.yaml:4-7); ``optimizer``: 'adam' or a dict of
        {learning_rate, beta_1, beta_2, epsilon} (engine.py:276-284)."""
        self.loss = L.get(loss) if loss is not None else L.WeightedCrossentropy()
        if isinstance(optimizer, str):
            if optimizer.lower() != 'adam':
                raise NotImplementedError(f'optimizer {optimizer!r}: the reference configs use adam only')
            optimizer = {}
        self.optimizer = dict(learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7)
        self.optimizer.update(optimizer or {})
        self.metrics = list(metrics or [])      # keras-style specs (metrics.yaml:1-23) or utils.metrics.Metric instances
        self._metric_set = None                 # utils.metrics.MetricSet, created on first use (needs the device)
        self._hyper_dirty = True
        self._invalidate_graphs()               # the loss configuration is baked into captured launch sequences

    def release_graphs(self):
        """Destroys every captured CUDA graph (they are re-captured on demand).  Call before
        ``torch.distributed.destroy_process_group()``: NCCL does not tear a communicator down while a live graph still
        holds collectives captured on it (the destroy call blocks)."""
        if self.params.device is not None:
            torch.cuda.synchronize()
        self._invalidate_graphs()
        if self.params.device is not None:
            torch.cuda.synchronize()

    def close(self):
        """Releases captured graphs and peer-memory mappings (call before destroying the process group)."""
        self.release_graphs()
        if getattr(self, '_p2p', None) is not None:
            self._p2p.check()
            self._p2p.close()

    def _invalidate_graphs(self):
        """Captured CUDA graphs hold the loss configuration (by value), the replica count and the bucket schedule:
        re-compiling or changing the data-parallel setup drops them (the next steps warm up and re-capture)."""
        for plan in self._plans.values():
            plan.graphs.clear()

    def metric_set(self):
        """The compiled pixel-threshold metrics (engine.py:273) as device counters, or None."""
        if self._metric_set is None and getattr(self, 'metrics', None):
            from .utils.metrics import MetricSet
            self._metric_set = MetricSet(self.metrics, self.device)
        return self._metric_set if self._metric_set else None

    def reset_metrics(self):
        ms = self.metric_set()
        if ms:
            ms.reset_state()

    def count_params(self, trainable=None):
        return self.params.count(trainable)

    def get_weights(self):
        return self.params.get_weights()

    def set_weights(self, weights):
        self.params.set_weights(weights)
        self._sync_replicas()

    def get_grads(self):
        return self.params.get_grads()

    # checkpoints: own format (TF checkpoints need TF), same ckpt-<step> naming as engine.py:52
    def save_weights(self, path, save_format=None, write=True):
        """Own ``.npz`` format by default; ``save_format='tf'`` writes a TensorFlow object-based checkpoint
        (``<path>.index`` + ``<path>.data-00000-of-00001``) laid out like the file the reference's
        ``ModelCheckpoint(save_weights_only=True)`` writes (engine.py:105), see ``utils/tf_checkpoint.py``.
        Under data parallelism EVERY rank calls it (the BatchNorm moving statistics are averaged over the replicas, a
        collective); ``write=False`` on the ranks that must not touch the file system."""
        if self._dp is not None and self.params.device is not None and self._dp.world_size > 1:
            self._dp.average(self.params.state)     # BN moving statistics are rank-local: saved as the replica mean
        if not write:
            return None
        os.makedirs(os.path.dirname(os.path.abspath(path)) or '.', exist_ok=True)
        if save_format == 'tf':
            from .utils import tf_checkpoint
            return tf_checkpoint.export_from(self, path)
        w = self.get_weights()
        extra = {}
        if self.params.device is not None:
            extra = {'__adam_m': self.params.m.cpu().numpy(), '__adam_v': self.params.v.cpu().numpy(),
                     '__step': self.params.step.cpu().numpy()}
        np.savez(path + '.npz' if not path.endswith('.npz') else path, **w, **extra)

    def load_weights(self, path):
        """engine.py:75,197,230.  Returns a status object with ``assert_existing_objects_matched()`` /
        ``assert_consumed()`` / ``expect_partial()`` like TF's checkpoint loader (engine.py:75 chains the first)."""
        tf_prefix = path[:-len('.index')] if path.endswith('.index') else path
        if os.path.exists(tf_prefix + '.index') and not os.path.exists(tf_prefix + '.npz'):
            # a TensorFlow checkpoint written by the reference (engine.py:105): read without TensorFlow
            from .utils import tf_checkpoint
            loaded, missing, unused = tf_checkpoint.load_into(self, tf_prefix)
            self._sync_replicas()
            st = LoadStatus(loaded + missing, loaded)
            st.unused = unused
            return st
        p = path if path.endswith('.npz') else path + '.npz'
        if os.path.isdir(path) and os.path.exists(os.path.join(path, 'weights.npz')):     # a directory written by save()
            p = os.path.join(path, 'weights.npz')
        z = np.load(p)
        names = [k for k in z.files if not k.startswith('__')]
        known = [k for k in names if k in self.params.specs]
        self.set_weights({k: z[k] for k in known})
        if '__step' in z.files and self.params.device is not None and z['__adam_m'].shape == tuple(self.params.m.shape):
            self.params.m.copy_(torch.from_numpy(z['__adam_m']))
            self.params.v.copy_(torch.from_numpy(z['__adam_v']))
            self.params.step.copy_(torch.from_numpy(z['__step']))
        self._sync_replicas()
        return LoadStatus(list(self.params.specs), names)

    def save(self, path, **kargs):
        """engine.py:226 ``model.save(path)``: a directory holding the model's class name + constructor config
        (``config.json``) and every variable incl. the optimizer slots (``weights.npz``); ``load_model`` rebuilds it."""
        import json
        os.makedirs(path, exist_ok=True)
        cfg = dict(class_name=type(self).__name__, config=self.get_config() if hasattr(self, 'get_config') else {},
                   input_shape=list(self.input_shape) if self.input_shape else None,
                   compute_dtype='bf16' if self.compute_dtype == torch.bfloat16 else 'fp32',
                   loss=dict(class_name='WeightedCrossentropy', config=self.loss.get_config()) if self.loss else None,
                   optimizer=self.optimizer)
        with open(os.path.join(path, 'config.json'), 'w') as f:
            json.dump(cfg, f, indent=1, default=str)
        self.save_weights(os.path.join(path, 'weights'))
        return path

    def _sync_replicas(self):
        """Mirrored variables (engine.py:260-263): after anything that rewrites variables on one replica
        (set_weights / load_weights / enabling data parallelism) rank 0's copy is broadcast to all."""
        if self._dp is not None and self.params.device is not None and self._dp.world_size > 1:
            ps = self.params
            self._dp.broadcast_parameters(ps.params, ps.state, ps.m, ps.v, ps.step)

    @staticmethod
    def list_checkpoints(save_dir):
        """engine.py:55-65: {step: path} of ``checkpoints/ckpt-<step>`` files."""
        out = {}
        d = os.path.join(save_dir, 'checkpoints')
        if os.path.isdir(d):
            for f in os.listdir(d):
                m = re.fullmatch(r'ckpt-(\d+)\.(npz|index)', f)       # own format, or the reference's TF checkpoints
                if m:
                    out[int(m.group(1))] = os.path.join(d, f[:-len(m.group(2)) - 1])
        return OrderedDict(sorted(out.items()))

    # ---- plans -----------------------------------------------------------------
    def _plan(self, batch, height, width, want_input_grad=False):
        key = (batch, height, width) if not want_input_grad else (batch, height, width, 'dx')
        if key not in self._plans:
            if not self.built:
                self.build((None, height, width, self.input_shape[-1] if self.input_shape else None))
            fresh = self.params.device is None
            self.params.materialize(self.device)
            if fresh:
                self._setup_p2p()
                self._sync_replicas()
            plan = R.Plan(self.params, batch, height, width, self.input_shape[-1], self.compute_dtype, self.device,
                          want_input_grad=want_input_grad)
            self._emit(plan)
            if os.environ.get('DNNCA_BN_FOLD', '1') != '0' and not want_input_grad:
                plan.fold_batchnorms()
            plan.fuse_bn_reductions()
            self._plans[key] = plan
        return self._plans[key]

    def training_plan(self, batch, height, width):
        """The plan ``train_step`` / ``forward_backward`` run for this input shape (MultiResUnet keeps a separate,
        channel-padded one beside its BatchNorm-folded inference plan)."""
        return self._plan(batch, height, width)

    def _run(self, plan, key, fn):
        """Runs ``fn`` (a fixed launch sequence) eagerly twice (warm-up: lazy attribute setup,
        allocator) and from then on as a CUDA graph replay."""
        if not self.use_cuda_graph:
            fn()
            return
        g = plan.graphs.get(key)
        if g is None:
            cnt = plan.graphs.get((key, 'warm'), 0)
            if cnt < 2:
                fn()
                plan.graphs[(key, 'warm')] = cnt + 1
                return
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            plan.graphs[key] = g
        g.replay()

    def _before_inference(self, plan):
        """Hook run with the batch already in ``plan.x_in``, before an inference launch sequence (models that cache
        folded / packed weights refresh them here)."""

    # ---- inference -------------------------------------------------------------
    def __call__(self, x, training=False):
        """``model(x)`` -> probabilities ``[B,H,W,1]`` (fp32 torch tensor on the device);
        the logits of the same call are kept in ``self.last_logits`` (keras caches them as
        ``_keras_logits``, losses.py:61)."""
        x = _to_device_unit(x, self.device)
        plan = self._plan(*x.shape[:3])
        plan.allocate(bool(training))
        plan.prestaged = False
        plan.x_in.copy_(x, non_blocking=True)
        self._before_inference(plan)
        if training:
            # keras ``model(x, training=True)``: BatchNormalization normalises with the batch statistics and
            # updates its moving averages; no gradients, no optimizer
            if not self.trainable_model:
                raise NotImplementedError(f'{type(self).__name__} is forward/inference-only in this build')

            def seq():
                plan.zero_step_state()
                plan.forward(True)
                plan.head_forward()
            self._run(plan, 'call_train', seq)
        else:
            def seq():
                plan.forward(False)
                plan.head_forward()
            self._run(plan, 'infer', seq)
        self.last_logits = plan.logits
        return plan.probs

    def input_gradient(self, x):
        """d(sum of output probabilities) / d(input), the quantity ``callbacks.py:290-299`` takes with a
        ``GradientTape`` around ``self.model(batch['x'])`` for its sensitivity maps.  Inference-mode forward
        (BatchNormalization with its moving statistics, so its backward is the per-channel scale), then the
        dgrad chain down to the network input (the first layer's dgrad, which training never needs).
        Returns ``(probs [B,H,W,1], dx [B,H,W,C])`` as fp32 device tensors."""
        if not self.trainable_model:
            raise NotImplementedError(f'{type(self).__name__} has no backward pass in this build')
        x = _to_device_unit(x, self.device)
        plan = self._plan(*x.shape[:3], want_input_grad=True)
        plan.allocate(True)
        plan.prestaged = False
        plan.x_in.copy_(x, non_blocking=True)

        def seq():
            plan.forward(False)
            plan.head_forward()
            plan.head_input_grad()
            plan.backward_inputs()
        self._run(plan, 'input_grad', seq)
        self.last_logits = plan.logits
        return plan.probs, plan.input_grad_f32()

    def predict(self, x, batch_size=None, verbose=0):
        if isinstance(x, (np.ndarray, torch.Tensor)):
            bs = batch_size or x.shape[0]
            outs = [self(x[i:i + bs]).cpu().numpy().copy() for i in range(0, x.shape[0], bs)]
            return np.concatenate(outs, 0)
        outs = []
        for batch in x:
            xb = batch[0] if isinstance(batch, (tuple, list)) else batch.get('x', batch)
            outs.append(self(xb).cpu().numpy().copy())
        return np.concatenate(outs, 0)

    # ---- training --------------------------------------------------------------
    def enable_data_parallel(self, process_group=None, bucket_bytes=8 << 20):
        """engine.py:260-263: synchronous data parallelism.  One process per GPU; gradients are
        SUM-all-reduced over NCCL with the loss pre-scaled by 1/world (== averaging)."""
        from . import parallel
        self._dp = parallel.GradAllReduce(process_group, bucket_bytes)
        self._p2p = None
        if self.built:
            self.params.materialize(self.device)
            self._setup_p2p()
        self._sync_replicas()                   # params, BN state, Adam slots and step counter from rank 0
        self._invalidate_graphs()
        return self

    def _setup_p2p(self):
        """Small models (whole gradient <= 8 MB: unet.yaml, mulmo_unet.yaml) exchange gradients through NVLink peer memory inside the Adam kernel
        (csrc/p2p_adam.cu) instead of NCCL; larger ones keep the bucketed NCCL all-reduce overlapped with backward."""
        from . import parallel
        ps = self.params
        if self._dp is None or getattr(self, '_p2p', None) is not None or ps.device is None:
            return
        n = max(ps.n_trainable, 4) + 4
        if not parallel.P2PAdam.usable(n * 4, self._dp.world_size):
            return
        p2p = parallel.P2PAdam(self._dp.group)
        full = p2p.setup(n, self.device)
        ps.rebind_grads(full, p2p.reduced)
        self._p2p = p2p
        self._invalidate_graphs()

    def _loss_cfg(self, plan):
        world = self._dp.world_size if self._dp else 1
        cfg = self.loss.native_config(plan.batch * plan.height * plan.width * world)
        return cfg

    def _sync_hyper(self, lr=None):
        if lr is not None and lr != self.optimizer['learning_rate']:
            self.optimizer['learning_rate'] = float(lr)
            self._hyper_dirty = True
        if getattr(self, '_hyper_dirty', True):
            o = self.optimizer
            self.params.hyper.copy_(torch.tensor([o['learning_rate'], o['beta_1'], o['beta_2'], o['epsilon']],
                                                 dtype=torch.float32), non_blocking=True)
            self._hyper_dirty = False

    def _side_stream(self):
        """Second stream of the training step (weight gradients beside the dgrad chain), or None (DNNCA_WGRAD_STREAM=0)."""
        # measured (B200, profiles/r02i_*): mulmo_unet +13 %, unet_big +2.8 %, unet.yaml -0.5 % -- the few-channel row kernels
        # are persistent one-CTA-per-SM kernels chained by programmatic dependent launch, which a second stream breaks
        mode = os.environ.get('DNNCA_WGRAD_STREAM', 'auto')
        if mode == '0' or (mode == 'auto' and self.params.n_trainable < 100_000):
            return None
        if getattr(self, '_side', None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def _adam(self):
        ps = self.params
        ps.version += 1
        N.call('dnnca_adam_step', N.stream_ptr(), N.ptr(ps.params), N.ptr(ps.grads), N.ptr(ps.m), N.ptr(ps.v),
               ps.n_trainable, N.ptr(ps.hyper), N.ptr(ps.step), N.ptr(ps.l2))

    def _loss_total(self, plan):
        """Scalar loss of the step into the tail of the flat gradient buffer (zeroed with the gradients, all-reduced
        with them): mean per-sample loss + L2 terms at the weights of THIS forward pass, scaled by 1/replicas."""
        ps = self.params
        world = self._dp.world_size if self._dp else 1
        N.call('dnnca_loss_total', N.stream_ptr(), N.ptr(plan.per_sample), plan.batch, N.ptr(ps.params), N.ptr(ps.l2),
               ps.n_trainable, 1.0 / world, N.ptr(ps.loss_in))

    # ---- label / weight validation (losses.py:30, 91-99) ------------------------------------------
    def _validate_labels(self, plan):
        """The reference asserts 0 <= label <= 1 (losses.py:91-92), 0 <= positive rate <= 1 (:97-98) and weight >= 0
        (:30) inside the loss.  Here the label statistics are reduced on the device anyway; they are read back and
        checked on the eager warm-up passes of every launch sequence (i.e. for the first batches of every
        shape), and on every step when ``DNNCA_VALIDATE=always`` (costs a device sync per step)."""
        import ctypes as C
        ls = torch.zeros(16, dtype=torch.uint8, device=self.device)
        N.call('dnnca_label_stats_init', N.stream_ptr(), N.ptr(ls))
        N.call('dnnca_label_stats', N.stream_ptr(), N.ptr(plan.y_in), plan.y_in.numel(), N.ptr(ls))
        host = N.LabelStats.from_buffer_copy(ls.cpu().numpy().tobytes())
        ssum, mn, mx = C.c_double(), C.c_float(), C.c_float()
        N.lib().dnnca_label_stats_decode(C.byref(host), C.byref(ssum), C.byref(mn), C.byref(mx))
        if not (mx.value <= 1.0 and mn.value >= 0.0):
            raise ValueError(f'labels must lie in [0, 1] (losses.py:91-92): min {mn.value}, max {mx.value}; '
                             'uint8 labels are divided by 255 on the device, float labels are taken as given')
        r = ssum.value / max(plan.y_in.numel(), 1)
        w = float(self.loss.weight) if self.loss.weight is not None else (1.0 / r if r > 0 else 1.0)
        w = self.loss.weight_mul * w + self.loss.weight_add
        if not w >= 0.0:
            raise ValueError(f'loss weight must be >= 0 (losses.py:30), got {w}')

    def _maybe_validate(self, plan, key):
        if os.environ.get('DNNCA_VALIDATE', 'warmup') == 'never':
            return
        if os.environ.get('DNNCA_VALIDATE') == 'always' or not self.use_cuda_graph or plan.graphs.get(key) is None:
            self._validate_labels(plan)

    def forward_backward(self, x, y):
        """Forward + loss + backward WITHOUT the optimizer step: leaves the gradients in
        ``get_grads()`` (parity tests, gradient inspection).  Returns the per-sample loss ``[B]``."""
        if self.loss is None:
            self.compile()
        if not self.trainable_model:
            raise NotImplementedError(f'{type(self).__name__} is forward/inference-only in this build')
        x, y = _to_device_unit(x, self.device), _to_device_unit(y, self.device)
        plan = self._plan(*x.shape[:3])
        plan.allocate(True)
        plan.prestaged = False
        plan.x_in.copy_(x, non_blocking=True)
        plan.y_in.copy_(y, non_blocking=True)
        self._maybe_validate(plan, 'fwdbwd')
        if self.loss.label_smoothing:
            plan.y_in.copy_(self.loss.prepare_labels(y), non_blocking=True)
        cfg = self._loss_cfg(plan)
        plan._cfg = cfg   # keep the ctypes struct alive for graph capture

        def seq():
            plan.zero_step_state()
            plan.forward(True)
            plan.head_loss(cfg)
            self._loss_total(plan)
            plan.backward()
            plan.end_backward()
        self._run(plan, 'fwdbwd', seq)
        self.last_logits = plan.logits
        return plan.per_sample

    # ---- input staging: the H2D copy of batch i+1 overlaps the compute of batch i ------------------
    def prefetch(self, x, y):
        """Starts the host->device copy of the NEXT batch on a side stream (the reference's
        ``tf.data`` pipeline ends with ``prefetch``, data.py:110).  ``x``/``y`` may be float32 in [0,1]
        (the reference's contract, data.py:193-206) or the raw uint8 slices (the /255 then runs on the
        device: 4x fewer PCIe bytes).  The following ``train_step(x, y)`` with the same objects uses it."""
        xt = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x))
        yt = _label_tensor(y)
        plan = self._plan(*xt.shape[:3])
        plan.allocate(True)
        st = getattr(plan, '_stage', None)
        if st is None or st['x'].dtype != xt.dtype or st['y'].dtype != yt.dtype or st['y'].shape != yt.shape:
            st = plan._stage = dict(x=torch.empty(xt.shape, dtype=xt.dtype, device=self.device),
                                    y=torch.empty(yt.shape, dtype=yt.dtype, device=self.device),
                                    stream=torch.cuda.Stream(), ready=torch.cuda.Event(), free=torch.cuda.Event(),
                                    key=None)
            st['free'].record(torch.cuda.current_stream())
        with torch.cuda.stream(st['stream']):
            st['stream'].wait_event(st['free'])          # previous consumer has copied the staging buffers out
            st['x'].copy_(xt, non_blocking=True)
            st['y'].copy_(yt, non_blocking=True)
            st['ready'].record(st['stream'])
        st['key'] = (id(x), id(y))
        return plan

    def _load_batch(self, plan, x, y):
        """Brings (x, y) into the plan's static input buffers, from the prefetch staging if it holds them."""
        st = getattr(plan, '_stage', None)
        if st is not None and st['key'] == (id(x), id(y)):
            cur = torch.cuda.current_stream()
            cur.wait_event(st['ready'])
            xs, ys = st['x'], st['y']
            st['key'] = None
        else:
            st = None
            xs = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x))
            ys = _label_tensor(y)
            xs, ys = xs.to(self.device, non_blocking=True), ys.to(self.device, non_blocking=True)
        packed = _is_packed(y)
        # uint8 slices of a bf16 plan go straight into the cast input buffer (one pass: /255 and rounding, data.py:206);
        # the graph variant keyed '..._u8' then skips the fp32 -> bf16 convert
        cast = getattr(plan, 'input_cast', None)
        plan.prestaged = bool(xs.dtype == torch.uint8 and cast is not None and cast.buf.data is not None and
                              cast.buf.data.dtype == torch.bfloat16)
        for src, dst in ((xs, plan.x_in), (ys, plan.y_in)):
            if dst is plan.y_in and packed:                                # binary masks shipped as bits (data_tail.pack_labels)
                assert src.numel() * 8 == dst.numel(), 'packed labels do not match the batch shape'
                N.call('dnnca_unpack_label_bits', N.stream_ptr(), N.ptr(src), dst.numel(), N.ptr(dst))
            elif src.dtype == torch.uint8:                                 # data.py:206: float32(uint8) / 255 on device
                if dst is plan.x_in and plan.prestaged:
                    N.call('dnnca_u8_to_unit', N.stream_ptr(), N.ptr(src), src.numel(), N.ptr(cast.buf.data), N.BF16)
                else:
                    N.call('dnnca_u8_to_unit', N.stream_ptr(), N.ptr(src), src.numel(), N.ptr(dst), N.F32)
            else:
                dst.copy_(src, non_blocking=True)
        if st is not None:
            st['free'].record(torch.cuda.current_stream())

    def train_step(self, x, y, lr=None):
        """One optimizer step (keras ``Model.train_step``): returns the scalar loss as a device
        tensor (mean per-sample loss + L2 regulariser terms; the global mean under data parallelism)."""
        if self.loss is None:
            self.compile()
        if not self.trainable_model:
            raise NotImplementedError(f'{type(self).__name__} is forward/inference-only in this build')
        shape = x.shape
        plan = self._plan(*shape[:3])
        plan.allocate(True)
        self._sync_hyper(lr)
        self._load_batch(plan, x, y)
        self._maybe_validate(plan, self._train_key(plan))
        loss = self._train_on_static(plan)
        ms = self.metric_set()
        if ms:                                   # keras updates the compiled metrics inside every train step
            ms.update_state(plan.y_metric if self.loss.label_smoothing else plan.y_in, plan.probs)
        return loss

    def _train_key(self, plan):
        return 'train_u8' if getattr(plan, 'prestaged', False) else 'train'

    def _train_on_static(self, plan):
        """The hot path once the batch sits in ``plan.x_in`` / ``plan.y_in``: ONE launch sequence = one CUDA graph
        {zero gradients + statistics, [label smoothing], forward, label statistics, fused head + loss, loss scalar,
        backward with the gradient buckets all-reduced as they complete, join, fused Adam}."""
        cfg = self._loss_cfg(plan)
        plan._cfg = cfg
        dp = self._dp if (self._dp is not None and self._dp.world_size > 1) else None
        p2p = getattr(self, '_p2p', None) if dp is not None else None
        ps = self.params
        if self.loss.label_smoothing and getattr(plan, 'y_metric', None) is None:
            plan.y_metric = torch.empty_like(plan.y_in)      # raw labels for the metrics (keras: the loss alone smooths)
            plan.y_tmp = torch.empty_like(plan.y_in)

        def seq():
            if p2p is not None:
                p2p.wait_done()                  # no peer may still be reading the gradients the next line zeroes
            plan.zero_step_state()
            if self.loss.label_smoothing:        # losses.py:62-67 on the device, inside the step's launch sequence
                plan.y_metric.copy_(plan.y_in)       # device-to-device copy node of the same graph
                N.call('dnnca_gaussian_filter2d', N.stream_ptr(), N.ptr(plan.y_metric), plan.batch, plan.height,
                       plan.width, int(self.loss.label_smoothing_filter_size), float(self.loss.label_smoothing_sigma),
                       N.ptr(plan.y_tmp), N.ptr(plan.y_in))
            plan.forward(True)
            plan.head_loss(cfg)
            self._loss_total(plan)
            if dp is None:
                plan.backward(side_stream=self._side_stream())
                plan.end_backward()
                self._adam()
            elif p2p is not None:
                plan.backward(side_stream=self._side_stream())
                plan.end_backward()
                ps.version += 1
                p2p.adam_step(ps)                # all-reduce over NVLink peer memory fused into the Adam kernel
            else:
                ready = plan.ready_frontier()
                dp.begin(ps.grads_full)
                dp.launch_ready(plan.pending_before_backward)
                plan.backward(after_op=lambda i: dp.launch_ready(ready[i], before=plan.join_side_streams),
                              side_stream=self._side_stream())
                plan.join_side_streams()
                plan.end_backward()
                dp.finish()
                self._adam()
        key = self._train_key(plan)
        if dp is not None and os.environ.get('DNNCA_DP_GRAPH', '1') == '0':
            saved, self.use_cuda_graph = self.use_cuda_graph, False     # escape hatch: eager launches, NCCL uncaptured
            try:
                self._run(plan, key, seq)
            finally:
                self.use_cuda_graph = saved
        else:
            self._run(plan, key, seq)
        self.last_logits = plan.logits
        return ps.loss_slot[0].clone()     # the slot is re-zeroed by the next step; callers may read the loss a step late

    def fit(self, x=None, y=None, validation_data=None, callbacks=None, steps_per_epoch=None, epochs=1,
            validation_freq=1, initial_epoch=0, verbose=1, lr_schedule=None, **kargs):
        """``Model.fit`` as the reference drives it (engine.py:126-135): ``x`` is an iterable of
        ``(features, labels)`` batches, one "epoch" is ``steps_per_epoch`` optimizer steps.
        ``lr_schedule(epoch, current_lr)`` reproduces the LearningRateScheduler callback
        (engine.py:97-100)."""
        callbacks = list(callbacks or [])
        hist = History(self, dict(epochs=epochs, steps=steps_per_epoch, verbose=verbose))
        for cb in callbacks:
            if hasattr(cb, 'set_model'):
                cb.set_model(self)
            if hasattr(cb, 'on_train_begin'):
                cb.on_train_begin({})
        it = iter(x) if y is None else None
        self.stop_training = False

        def next_batch():
            nonlocal it
            if it is None:
                return x, y
            try:
                return next(it)
            except StopIteration:
                it = iter(x)
                return next(it)
        nxt = next_batch()
        self.prefetch(*nxt)
        for epoch in range(initial_epoch, epochs):
            lr = lr_schedule(epoch, self.optimizer['learning_rate']) if lr_schedule else None
            self.reset_metrics()
            losses = []
            n = steps_per_epoch or 1
            self._last_epoch_steps = n           # batches of this epoch (engine.ModelCheckpoint counts them, keras save_freq)
            for _ in range(n):
                xb, yb = nxt
                losses.append(self.train_step(xb, yb, lr=lr))
                nxt = next_batch()
                self.prefetch(*nxt)          # H2D of the next batch overlaps the step just launched
            logs = {'loss': float(torch.stack(losses).mean())}
            if self.metric_set():
                logs.update(self.metric_set().result())
            if validation_data is not None and validation_freq and (epoch + 1) % validation_freq == 0:
                logs.update({'val_' + k: v for k, v in self.evaluate(validation_data, return_dict=True).items()})
            hist.record(epoch, logs)
            for cb in callbacks:
                if hasattr(cb, 'on_epoch_end'):
                    cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        for cb in callbacks:
            if hasattr(cb, 'on_train_end'):
                cb.on_train_end({})
        return hist

    def evaluate(self, x, callbacks=None, verbose=0, return_dict=True):
        """engine.py:198-203: forward with training=False + loss over a dataset of (features, labels)."""
        if self.loss is None:
            self.compile()
        tot, cnt = 0.0, 0
        ms = self.metric_set()
        if ms:
            ms.reset_state()
        for xb, yb in x:
            xb, yb = _to_device_unit(xb, self.device), _to_device_unit(yb, self.device)
            plan = self._plan(*xb.shape[:3])
            plan.allocate(False)
            plan.prestaged = False
            plan.x_in.copy_(xb, non_blocking=True)
            plan.y_in.copy_(yb, non_blocking=True)
            self._before_inference(plan)
            self._maybe_validate(plan, 'eval')
            if self.loss.label_smoothing:
                plan.y_in.copy_(self.loss.prepare_labels(yb), non_blocking=True)
            cfg = self._loss_cfg(plan)
            plan._cfg_eval = cfg

            def seq():
                plan.forward(False)
                plan.head_loss(cfg, with_grads=False)
            self._run(plan, 'eval', seq)
            self.last_logits = plan.logits
            if ms:
                ms.update_state(yb if self.loss.label_smoothing else plan.y_in, plan.probs)
            tot += float(plan.per_sample.sum())
            cnt += plan.batch
        out = {'loss': tot / max(cnt, 1)}
        if ms:
            out.update(ms.result())
        return out if return_dict else out['loss']
