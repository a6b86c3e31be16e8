"""``Model.fit`` / ``evaluate`` driven the way the reference's engine drives them (engine.py:126-135:
``fit(ds, validation_data=, callbacks=, steps_per_epoch=1, epochs=max_steps, validation_freq=, initial_epoch=)`` with the
LearningRateScheduler of engine.py:97-100), against the oracle's training loop."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_models as rm
from oracle import ref_ops as ops
from tests.golden.make_golden import CASES

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class _Recorder:
    """keras.callbacks.Callback surface the engine's callbacks use (callbacks.py:28-52): set_model, on_train_begin,
    on_epoch_end(epoch, logs), on_train_end; may set model.stop_training."""

    def __init__(self, stop_at=None):
        self.events, self.stop_at, self.model = [], stop_at, None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        self.events.append('begin')

    def on_epoch_end(self, epoch, logs=None):
        self.events.append((epoch, dict(logs)))
        if self.stop_at is not None and epoch >= self.stop_at:
            self.model.stop_training = True

    def on_train_end(self, logs=None):
        self.events.append('end')


def _oracle_eval(ref, val, loss_cfg):
    tot = cnt = 0
    for xb, yb in val:
        out = ref.forward(xb, training=False)
        per = ops.weighted_crossentropy(torch.tensor(yb), out['logits'], **loss_cfg)
        tot += float(per.sum())
        cnt += len(per)
    return tot / cnt


def test_fit_follows_the_engine_loop_fp32():
    from dnncancerannotator_b200.models import tf_models
    from dnncancerannotator_b200.synthetic import make_slices
    case = 'unet_bn_tiny'
    model, opts, B, H, C, loss_cfg, _ = CASES[case]
    z = np.load(os.path.join(GOLDEN, case + '.npz'))
    weights = {k[2:]: z[k] for k in z.files if k.startswith('w:')}
    m = getattr(tf_models, model)(**opts, dtype='fp32')
    m.build((None, H, H, C))
    m.set_weights(weights)
    m.compile(optimizer='adam', loss=dict(class_name='WeightedCrossentropy', config=loss_cfg))
    ref = rm.build_model(model, opts, (None, H, H, C), seed=0)
    ref.set_weights(weights)
    train = [make_slices(B, H, H, C, seed=100 + i) for i in range(3)]          # the dataset repeats (data.py:108 `.repeat()`)
    val = [make_slices(2, H, H, C, seed=200 + i) for i in range(2)]
    sched = lambda epoch, lr: 1e-3 * 0.5 ** (epoch // 2)                        # engine.py:97-100 form: f(epoch, current_lr)
    rec = _Recorder()
    hist = m.fit(train, validation_data=val, callbacks=[rec], steps_per_epoch=1, epochs=6, validation_freq=2,
                 initial_epoch=1, verbose=0, lr_schedule=sched)
    # the oracle's loop: one optimizer step per "epoch", batches in dataset order, keras-form Adam, evaluation every 2nd
    mom = {k: (torch.zeros_like(ref.weights[k]), torch.zeros_like(ref.weights[k])) for k in ref.trainable}
    want_loss, want_val = [], []
    for i, epoch in enumerate(range(1, 6)):
        xb, yb = train[i % 3]
        r = ref.train_step_grads(xb, yb, loss_cfg)
        want_loss.append(r['loss'])
        for k in ref.trainable:
            ref.weights[k], m_, v_ = ops.adam_step(ref.weights[k], r['grads'][k], mom[k][0], mom[k][1], i + 1, lr=sched(epoch, None))
            mom[k] = (m_, v_)
        for k, v in r['new_moving'].items():
            ref.weights[k] = v
        if (epoch + 1) % 2 == 0:
            want_val.append(_oracle_eval(ref, val, loss_cfg))
    assert hist.epoch == [1, 2, 3, 4, 5] and hist.model is m and hist.params['epochs'] == 6
    np.testing.assert_allclose(hist.history['loss'], want_loss, rtol=2e-3)
    np.testing.assert_allclose(hist.history['val_loss'], want_val, rtol=2e-3)
    assert len(want_val) == 3                                                    # epochs 1, 3, 5
    assert rec.events[0] == 'begin' and rec.events[-1] == 'end' and rec.model is m
    assert [e[0] for e in rec.events[1:-1]] == [1, 2, 3, 4, 5]
    assert 'val_loss' in rec.events[1][1] and 'val_loss' not in rec.events[2][1]
    # a callback that sets model.stop_training ends the loop after that epoch (keras EarlyStopping protocol)
    rec2 = _Recorder(stop_at=1)
    h2 = m.fit(train, callbacks=[rec2], steps_per_epoch=2, epochs=10, verbose=0)
    assert h2.epoch == [0, 1] and 'val_loss' not in h2.history


def test_multiresunet_fit_alternates_training_and_inference_plans_fp32():
    """fit() on MultiResUnet: training steps run the channel-padded plan, the validation pass after every epoch runs the
    BatchNorm-folded inference plan, which must re-fold the variables the step just updated."""
    from dnncancerannotator_b200.models import tf_models
    from dnncancerannotator_b200.synthetic import make_slices
    mo = dict(height=None, width=None, n_channels=5)
    ref = rm.build_model('MultiResUnet', mo, None, seed=0)
    ref.randomize_bn(seed=1)
    m = tf_models.MultiResUnet(**mo, dtype='fp32')
    m.build((None, 32, 32, 5))
    m.set_weights(ref.get_weights())
    loss_cfg = dict(weight_mul=3.0)
    m.compile(loss=dict(class_name='WeightedCrossentropy', config=loss_cfg))
    train = [make_slices(2, 32, 32, 5, seed=1235), make_slices(2, 32, 32, 5, seed=35)]
    val = [make_slices(2, 32, 32, 5, seed=23)]
    vals = []

    class _Check(_Recorder):
        def on_epoch_end(self, epoch, logs=None):
            ref.set_weights(self.model.get_weights())
            vals.append((logs['val_loss'], _oracle_eval(ref, val, loss_cfg)))
    hist = m.fit(train, validation_data=val, callbacks=[_Check()], steps_per_epoch=2, epochs=4, verbose=0)
    assert len(vals) == 4 and np.isfinite(hist.history['loss']).all()
    for got, want in vals:                     # every epoch: the inference plan evaluated the CURRENT variables
        assert abs(got - want) <= 1e-4 * abs(want), (got, want)
    assert len({round(v[0], 6) for v in vals}) == 4            # and they did change from epoch to epoch


def _engine_cfg(model='UNetAnnotator', **model_options):
    return dict(model=model, model_options=model_options,
                deploy_options=dict(optimizer='adam', LearningRateScheduler='lambda epoch, current_lr: 0.001 * 0.9 ** (epoch // 2)',
                                    loss=dict(class_name='WeightedCrossentropy', config=dict(weight_mul=3.0)),
                                    enable_multigpu=False))


def test_engine_train_checkpoints_resume_and_eval(tmp_path):
    """engine.py:80-137 / 139-210 through ``dnncancerannotator_b200.engine.TFKerasModel``: ``checkpoints/ckpt-<step>`` every
    ``save_freq`` steps (TensorFlow-format files), validation at the same frequency, auto-resume from the latest
    checkpoint (variables, Adam slots, iteration count, LR-schedule position) reproducing the uninterrupted run,
    evaluation over the checkpoints with ``step_range`` / ``min_interval`` / ``results.csv``."""
    from dnncancerannotator_b200 import engine as E
    from dnncancerannotator_b200.synthetic import make_slices
    cfg = _engine_cfg(n_filters_first=4, n_downsample=2, rate=2, kernel_size=3, conv_stride=1, bn=True, padding='same')
    train = [make_slices(3, 32, 32, 3, seed=100 + i) for i in range(3)]
    val = [make_slices(2, 32, 32, 3, seed=200)]
    sp, sp_full = str(tmp_path / 'run'), str(tmp_path / 'run_full')
    eng = E.TFKerasModel(cfg, dtype='fp32')
    h1 = eng.train(train, val_data=val, save_path=sp, save_freq=3, max_steps=7)
    assert h1.epoch == list(range(7)) and len(h1.history['val_loss']) == 2          # validated after steps 3 and 6
    ck = eng.get_ckpts(os.path.join(sp, 'checkpoints'))
    assert list(ck) == [3, 6] and all(os.path.exists(p + '.index') and os.path.exists(p + '.data-00000-of-00001') for p in ck.values())
    full = E.TFKerasModel(cfg, dtype='fp32')
    hf = full.train(train, val_data=val, save_path=sp_full, save_freq=3, max_steps=10)
    np.testing.assert_allclose(hf.history['loss'][:7], h1.history['loss'], rtol=1e-4)
    # a fresh process resumes: latest checkpoint, step counter -> initial_epoch
    eng2 = E.TFKerasModel(cfg, dtype='fp32')
    with pytest.warns(UserWarning, match='Resumed from 6'):
        h2 = eng2.train(train, val_data=val, save_path=sp, save_freq=3, max_steps=10)
    assert eng2.current_step == 6 and h2.epoch == [6, 7, 8, 9]
    np.testing.assert_allclose(h2.history['loss'], hf.history['loss'][6:], rtol=2e-4)
    np.testing.assert_allclose(h2.history['val_loss'], hf.history['val_loss'][2:], rtol=2e-4)
    assert list(eng2.get_ckpts(os.path.join(sp, 'checkpoints'))) == [3, 6, 9]
    # evaluation over the checkpoints
    res = eng2.eval(val, sp, tag='val', export_csv=True)
    assert list(res) == [3, 6, 9] and abs(res[9]['loss'] - h2.history['val_loss'][-1]) <= 1e-6 * abs(res[9]['loss'])
    assert abs(res[6]['loss'] - h1.history['val_loss'][-1]) <= 1e-6 * abs(res[6]['loss'])
    csv = open(os.path.join(sp, 'tfevents', 'val', 'results.csv')).read().splitlines()
    assert csv[0].startswith('step') and [int(l.split(',')[0]) for l in csv[1:]] == [3, 6, 9]
    with pytest.raises(ValueError):
        eng2.eval(val, sp, tag='val')                                                 # "tag: val already exists."
    assert list(eng2.eval(val, sp, tag='val', avoid_overwrite=True, step_range=(4, 9))) == [6, 9]
    with pytest.warns(UserWarning, match='min_interval'):
        assert list(eng2.eval(val, sp, tag='other', min_interval=4)) == [3, 9]
    assert eng2.predict(val[0][0]).shape == (2, 32, 32, 1)


def test_engine_multiresunet_checkpoints_and_resume(tmp_path):
    from dnncancerannotator_b200 import engine as E
    from dnncancerannotator_b200.synthetic import make_slices
    cfg = _engine_cfg('MultiResUnet', height=None, width=None, n_channels=5)
    train = [make_slices(2, 32, 32, 5, seed=1235)]
    sp = str(tmp_path / 'run')
    eng = E.TFKerasModel(cfg, dtype='fp32')
    h1 = eng.train(train, val_data=train, save_path=sp, save_freq=1, max_steps=2)
    ck = eng.get_ckpts(os.path.join(sp, 'checkpoints'))
    assert list(ck) == [1, 2] and all(os.path.exists(p + '.npz') for p in ck.values())     # own format
    eng2 = E.TFKerasModel(cfg, dtype='fp32')
    with pytest.warns(UserWarning, match='Resumed from 2'):
        h2 = eng2.train(train, val_data=train, save_path=sp, save_freq=1, max_steps=3)
    assert h2.epoch == [2] and np.isfinite(h2.history['loss'][0]) and h2.history['loss'][0] < h1.history['loss'][0]
    w1, w2 = eng.model.get_weights(), eng2.model.get_weights()
    assert any(not np.array_equal(w1[k], w2[k]) for k in w1)                # it trained on from the restored variables
