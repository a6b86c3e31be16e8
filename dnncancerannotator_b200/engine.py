"""Host-side mirror of the reference's ``annotator/engine.py`` ``TFKerasModel`` -- the immediate CALLER of the hot path:
model / loss / metrics / Adam wiring from the layered config (engine.py:254-288), the training loop with periodic
``checkpoints/ckpt-<step>`` files, auto-resume and early stopping (engine.py:80-137), evaluation over the saved
checkpoints (engine.py:139-210) and checkpoint discovery (engine.py:52-78, 212-220).

It adds no arithmetic: everything below calls ``keras_like.Model`` (``compile`` / ``fit`` / ``evaluate`` / ``predict`` /
``save_weights`` / ``load_weights``).  What the reference delegates to Keras callbacks is restated here with Keras's
semantics [TF-semantics]: ``ModelCheckpoint(filepath, save_freq=<int>, save_weights_only=True)`` saves after every
``save_freq`` batches seen in this ``fit`` call, formatting ``{epoch}`` with the 1-based epoch; ``EarlyStopping(patience)``
monitors ``val_loss`` (minimum, ``min_delta`` 0), ignores epochs whose logs hold no ``val_loss`` and stops once ``patience``
monitored epochs in a row did not improve; ``LearningRateScheduler(schedule)`` sets ``lr = schedule(epoch, lr)`` at the start
of every epoch.  TensorBoard, the Visualizer and the tqdm progress callback (engine.py:108-123) are out of scope.

Multi-GPU: the reference's ``MirroredStrategy`` (engine.py:260-263, ``deploy_options.enable_multigpu``) is one process driving
all GPUs; here it is one process per GPU (``torchrun``): with ``enable_multigpu`` and an initialised ``torch.distributed``
group of more than one rank the model runs synchronous data parallelism (``Model.enable_data_parallel``); checkpoints are
written by rank 0.
"""
from __future__ import annotations

import copy
import os
import re
import warnings
from collections import OrderedDict

import numpy as np

from .models import tf_models


# ---- Keras callbacks the engine uses ------------------------------------------------------------------------------------
class Callback:
    model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass

    def on_train_end(self, logs=None):
        pass


class ModelCheckpoint(Callback):
    """``tf.keras.callbacks.ModelCheckpoint(filepath, save_freq=<int>, save_weights_only=True)`` (engine.py:105).
    ``save_format``: 'tf' writes the TensorFlow checkpoint the reference itself reads (``<path>.index`` + data shard);
    'npz' this package's own file; 'auto' = 'tf' where the variable naming of the reference is restated
    (UNetAnnotator / MulmoUNetAnnotator), else 'npz'."""

    def __init__(self, filepath, save_freq=100, save_weights_only=True, save_format='auto', is_writer=True):
        if not save_weights_only:
            raise NotImplementedError('ModelCheckpoint: the reference saves weights only (engine.py:105)')
        self.filepath, self.save_freq, self.save_format, self.is_writer = filepath, int(save_freq), save_format, is_writer
        self.seen = 0
        self.saved = []

    def on_train_begin(self, logs=None):
        self.seen = 0

    def on_batches(self, epoch, n_batches, logs=None):
        """``n_batches`` batches of ``epoch`` have just run (``fit`` reports them at the end of the epoch: with the
        engine's ``steps_per_epoch=1`` that is Keras's per-batch check)."""
        self.seen += n_batches
        if self.seen >= self.save_freq:
            self.seen = 0
            path = self.filepath.format(epoch=epoch + 1, **(logs or {}))
            self._save(path)                      # every rank takes part (replica mean of the BatchNorm statistics)
            if self.is_writer:
                self.saved.append(path)

    def _save(self, path):
        fmt = self.save_format
        if fmt == 'auto':
            fmt = 'tf' if type(self.model).__name__ in ('UNetAnnotator', 'MulmoUNetAnnotator') else 'npz'
        self.model.save_weights(path, save_format='tf' if fmt == 'tf' else None, write=self.is_writer)

    def on_epoch_end(self, epoch, logs=None):
        self.on_batches(epoch, self.model._last_epoch_steps, logs)


class EarlyStopping(Callback):
    """``tf.keras.callbacks.EarlyStopping(patience=...)`` with its defaults (monitor 'val_loss', min_delta 0, mode min)."""

    def __init__(self, patience=0, monitor='val_loss', min_delta=0.0, verbose=0):
        self.patience, self.monitor, self.min_delta = int(patience), monitor, abs(float(min_delta))
        self.wait, self.best, self.stopped_epoch = 0, np.inf, 0

    def on_train_begin(self, logs=None):
        self.wait, self.best, self.stopped_epoch = 0, np.inf, 0

    def on_epoch_end(self, epoch, logs=None):
        current = (logs or {}).get(self.monitor)
        if current is None:
            return                      # Keras warns and skips: with validation_freq > 1 most epochs have no val_loss
        self.wait += 1
        if current < self.best - self.min_delta:
            self.best, self.wait = current, 0
        if self.wait >= self.patience and epoch > 0:
            self.stopped_epoch = epoch
            self.model.stop_training = True


def solve_learning_rate_scheduler(spec):
    """``eval(self.learning_rate_scheduler)`` of engine.py:97-100: the config holds the source of a
    ``lambda epoch, current_lr: ...`` (deploy_options.yaml:3, lrdecay_high_init.yaml:1-2)."""
    if spec is None or callable(spec):
        return spec
    import math
    names = dict(math=math, min=min, max=max, abs=abs, pow=pow, round=round, float=float, int=int)
    fn = eval(spec, {'__builtins__': {}}, names)     # noqa: S307 -- the reference evaluates the same config string (plain eval)
    if not callable(fn):
        raise ValueError(f'LearningRateScheduler must evaluate to a callable, got {spec!r}')
    return fn


# ---- the engine ----------------------------------------------------------------------------------------------------------
class TFKerasModel:
    """``annotator.engine.TFKerasModel`` (engine.py:36-288) over the B200 models."""

    ckpt_pattern = 'ckpt-{epoch}'

    def __init__(self, model_config, dtype=None):
        self.model_config = copy.deepcopy(model_config)
        self._dtype = dtype
        self.model = self.from_config(model_config)
        self.current_step = 0

    # engine.py:254-288
    def from_config(self, model_config):
        assert 'model' in model_config
        assert 'model_options' in model_config
        assert 'deploy_options' in model_config
        deploy_options = copy.deepcopy(model_config['deploy_options'])
        self.enable_multigpu = deploy_options.pop('enable_multigpu', True)
        self.learning_rate_scheduler = deploy_options.pop('LearningRateScheduler', None)
        model = getattr(tf_models, model_config['model'])(**model_config['model_options'], dtype=self._dtype)
        model.compile(optimizer=deploy_options.get('optimizer', 'adam'), loss=deploy_options.get('loss'),
                      metrics=list(deploy_options.get('metrics', [])))
        return model

    def _world(self):
        import torch.distributed as dist
        if self.enable_multigpu and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.get_world_size(), dist.get_rank()
        return 1, 0

    def _enter_strategy_section(self):
        world, _ = self._world()
        if world > 1 and self.model._dp is None:
            self.model.enable_data_parallel()

    @staticmethod
    def _peek_input_shape(dataset):
        """``dataset.element_spec[0].shape`` of the reference (engine.py:93): read from the first batch of a re-iterable
        dataset; a one-shot iterator would lose that batch, so it needs ``input_shape=``."""
        if iter(dataset) is dataset:
            raise ValueError('a one-shot iterator cannot be inspected without consuming a batch: pass input_shape=(None, H, W, C)')
        first = next(iter(dataset))[0]
        return (None, *first.shape[1:])

    # engine.py:55-78
    def get_ckpts(self, base_path):
        """{step: path prefix} of ``ckpt-<step>`` files (TensorFlow ``.index`` checkpoints as the reference lists them, and
        this package's ``.npz`` files), ascending."""
        out = {}
        if os.path.isdir(base_path):
            pat = self.ckpt_pattern.format(epoch=r'(\d+)')
            for f in os.listdir(base_path):
                m = re.fullmatch(pat + r'\.(index|npz)', f)
                if m:
                    out[int(m.group(1))] = os.path.join(base_path, f[:-len(m.group(2)) - 1])
        return OrderedDict(sorted(out.items()))

    def _auto_resume(self, base_path):
        ckpts = self.get_ckpts(base_path)
        if not ckpts:
            return
        latest_step = max(ckpts)
        self.model.load_weights(ckpts[latest_step]).assert_existing_objects_matched()
        self.current_step = latest_step
        warnings.warn(f'Resumed from {latest_step}')

    # engine.py:80-137
    def train(self, dataset, val_data=None, save_path=None, save_freq=100, max_steps=None, early_stop_steps=None,
              visualization=None, auto_resume=True, profile=False, input_shape=None):
        """``dataset``: iterable of ``(features, labels)`` batches (it is cycled, like ``.repeat()`` data.py:108).
        ``input_shape`` replaces ``dataset.element_spec[0].shape`` when the dataset is a plain list / generator."""
        if visualization:
            raise NotImplementedError('the Visualizer callback (callbacks.py:55-446) is outside the hot path')
        self._enter_strategy_section()
        if not self.model.built:
            self.model.build(tuple(input_shape or self._peek_input_shape(dataset)))
        if self.model.params.device is None:
            self.model.params.materialize(self.model.device)      # Adam slots / step counter can be restored
        if auto_resume and save_path is not None:
            self._auto_resume(os.path.join(save_path, 'checkpoints'))
        callbacks = []
        if save_path is not None:
            ckpt_path = os.path.join(save_path, 'checkpoints', self.ckpt_pattern)
            os.makedirs(os.path.dirname(ckpt_path), exist_ok=True)
            callbacks.append(ModelCheckpoint(ckpt_path, save_freq=save_freq, save_weights_only=True, is_writer=self._world()[1] == 0))
        if early_stop_steps is not None:
            callbacks.append(EarlyStopping(patience=early_stop_steps, verbose=1))
        if max_steps is None:
            raise ValueError('max_steps is required (keras fit(epochs=None) fails the same way)')
        return self.model.fit(dataset, validation_data=val_data, callbacks=callbacks, steps_per_epoch=1, epochs=max_steps,
                              validation_freq=save_freq, initial_epoch=self.current_step, verbose=0,
                              lr_schedule=solve_learning_rate_scheduler(self.learning_rate_scheduler))

    # engine.py:139-210
    def eval(self, dataset, save_path, viz_ds=None, tag='val', avoid_overwrite=False, export_path=None, export_images=False,
             visualize_sensitivity=False, export_csv=False, min_interval=1, step_range=None, overlay=False,
             export_casewise_metrics=False, input_shape=None):
        """Evaluates every checkpoint under ``save_path/checkpoints`` (filtered by ``step_range`` / ``min_interval``) on
        ``dataset``; returns ``{step: results}`` and, with ``export_csv``, writes ``<export_path>/<tag>/results.csv`` like
        the reference (which returns nothing)."""
        if viz_ds is not None or export_images or visualize_sensitivity:
            raise NotImplementedError('the Visualizer callback (callbacks.py:55-446) is outside the hot path')
        self._enter_strategy_section()
        if not self.model.built:
            self.model.build(tuple(input_shape or self._peek_input_shape(dataset)))
        if self.model.params.device is None:
            self.model.params.materialize(self.model.device)
        ckpt_path = os.path.join(save_path, 'checkpoints')
        if not export_path:
            export_path = os.path.join(save_path, 'tfevents')
        if os.path.exists(os.path.join(export_path, tag)):
            if avoid_overwrite:
                while os.path.exists(os.path.join(export_path, tag)):
                    tag += '_'
            else:
                raise ValueError(f'tag: {tag} already exists.')
        if step_range is None:
            step_range = 0, float('inf')
        else:
            assert len(step_range) == 2
            assert 0 <= step_range[0] <= step_range[1]
        results, previous_step = OrderedDict(), None
        for ckpt_step, ckpt_path_ in self.get_ckpts(ckpt_path).items():
            if not step_range[0] <= ckpt_step <= step_range[1]:
                continue
            if previous_step is not None and (ckpt_step - previous_step) < min_interval:
                warnings.warn(f'Ignored {ckpt_path_} due to min_interval:{min_interval}.')
                continue
            previous_step = ckpt_step
            self.load(ckpt_path_)
            results[ckpt_step] = self.model.evaluate(dataset, callbacks=[], verbose=0, return_dict=True)
        if export_csv and self._world()[1] == 0:
            import pandas as pd
            os.makedirs(os.path.join(export_path, tag), exist_ok=True)
            frame = pd.DataFrame.from_dict(results, orient='index')
            frame.index.rename('step', inplace=True)
            frame.to_csv(os.path.join(export_path, tag, 'results.csv'))
        return results

    # engine.py:212-236
    def list_ckpts(self, save_path):
        assert os.path.exists(save_path)
        return self.get_ckpts(save_path)

    def predict(self, dataset):
        return self.model.predict(dataset)

    def save(self, path, fileformat=None):
        self.model.save(path)
        return self

    def load(self, path):
        self.model.load_weights(path)
        return self

    def get_config(self):
        return self.model_config
