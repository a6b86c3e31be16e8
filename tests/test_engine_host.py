"""Host logic of ``dnncancerannotator_b200.engine`` (mirror of the reference's ``annotator/engine.py``): config wiring for
every shipped model config x overlay, the Keras callback semantics it restates, checkpoint discovery.  No GPU."""
import glob
import os

import numpy as np
import pytest

from dnncancerannotator_b200 import engine as E
from dnncancerannotator_b200.utils.load import load_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, 'configs')
ADD = os.path.join(CFG, 'additionals')


def _cfg(model, *overlays):
    return load_config([os.path.join(CFG, model + '.yaml'), os.path.join(ADD, 'data_options.yaml'),
                        os.path.join(ADD, 'deploy_options.yaml')] + [os.path.join(ADD, o + '.yaml') for o in overlays])


@pytest.mark.parametrize('model', ['unet', 'unet_big', 'mulmo_unet', 'multiresunet'])
def test_from_config_every_model_config(model):
    """engine.py:254-288: model from the registry, loss from its dict, Adam, metrics list, multi-GPU flag."""
    cfg = _cfg(model, 'metrics')
    eng = E.TFKerasModel(cfg)
    m = eng.model
    assert type(m).__name__ == cfg['model'] and eng.get_config() == cfg
    assert eng.enable_multigpu is False                                   # deploy_options.yaml:8
    assert m.loss.weight_mul == 3.0 and m.optimizer['learning_rate'] == 1e-3 and m.optimizer['epsilon'] == 1e-7
    assert len(m.metrics) == len(cfg['deploy_options']['metrics']) >= 6
    sched = E.solve_learning_rate_scheduler(eng.learning_rate_scheduler)
    assert sched(0, None) == 1e-3 and abs(sched(2500, None) - 1e-3 * 0.96 ** 2) < 1e-12
    # the key is absent -> the reference's code default is multi-GPU ON (engine.py:259)
    cfg2 = _cfg(model)
    del cfg2['deploy_options']['enable_multigpu']
    assert E.TFKerasModel(cfg2).enable_multigpu is True


def test_every_overlay_parses_and_wires():
    """configs/additionals/*.yaml on top of unet.yaml: the dotted-key overlays land where engine.py / data.py read them."""
    names = sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(ADD, '*.yaml')))
    assert len(names) >= 20
    for o in names:
        cfg = _cfg('unet', o)
        eng = E.TFKerasModel(cfg)
        if o == 'leakyReLU':
            assert eng.model.configs['activation'] == dict(class_name='LeakyReLU', config=dict(alpha=0.3))
        if o == 'kernel_regularizer':
            assert eng.model.configs['kernel_regularizer']['config']['l2'] == 0.01
            assert any(s['l2'] == 0.01 for s in _built(eng).params.specs.values())
        if o == 'lrdecay_high_init':
            assert E.solve_learning_rate_scheduler(eng.learning_rate_scheduler)(1000, None) == pytest.approx(0.005 * 0.96)
        if o == 'multigpu':
            assert eng.enable_multigpu is True
        if o == 'enable_label_smoothing':
            assert eng.model.loss.label_smoothing
        if o == 'train_batch28':
            assert cfg['data_options']['train']['batch_size'] == 28
        if o.startswith('slice_type_'):
            assert cfg['data_options']['train']['slice_types'][-1] == 'label'
        if o == 'metrics':
            assert any('RegionBasedRecall' in d for d in cfg['deploy_options']['metrics'])


def _built(eng):
    eng.model.build((None, 32, 32, 3))
    return eng.model


class _FakeModel:
    def __init__(self):
        self.saved, self.stop_training, self._last_epoch_steps = [], False, 1

    def save_weights(self, path, save_format=None, write=True):
        self.saved.append((path, save_format, write))


def test_model_checkpoint_counts_batches_like_keras():
    """save_freq batches seen since the start of THIS fit; {epoch} is 1-based: with the engine's steps_per_epoch=1 and a
    resume at step 6, save_freq=3 writes ckpt-9, ckpt-12 ..."""
    m = _FakeModel()
    cb = E.ModelCheckpoint('/x/ckpt-{epoch}', save_freq=3)
    cb.set_model(m)
    cb.on_train_begin()
    for epoch in range(6, 13):
        cb.on_epoch_end(epoch, {'loss': 1.0})
    assert [p for p, _, _ in m.saved] == ['/x/ckpt-9', '/x/ckpt-12']
    assert cb.saved == ['/x/ckpt-9', '/x/ckpt-12']
    # a rank that does not write still takes part in the call (replica mean of the BatchNorm statistics is a collective)
    m2 = _FakeModel()
    cb2 = E.ModelCheckpoint('/x/ckpt-{epoch}', save_freq=2, is_writer=False)
    cb2.set_model(m2)
    cb2.on_train_begin()
    for epoch in range(4):
        cb2.on_epoch_end(epoch, {})
    assert [(p, w) for p, _, w in m2.saved] == [('/x/ckpt-2', False), ('/x/ckpt-4', False)] and cb2.saved == []
    with pytest.raises(NotImplementedError):
        E.ModelCheckpoint('/x', save_weights_only=False)


def test_early_stopping_follows_keras():
    m = _FakeModel()
    cb = E.EarlyStopping(patience=2)
    cb.set_model(m)
    cb.on_train_begin()
    vals = {0: 1.0, 1: None, 2: 0.9, 3: None, 4: 0.95, 5: None, 6: 0.91, 7: 0.5}
    stopped = None
    for epoch, v in vals.items():
        cb.on_epoch_end(epoch, {'loss': 0.1} if v is None else {'loss': 0.1, 'val_loss': v})
        if m.stop_training:
            stopped = epoch
            break
    assert stopped == 6 and cb.best == 0.9 and cb.stopped_epoch == 6        # two monitored epochs without improvement


def test_get_ckpts_lists_tf_and_npz_checkpoints(tmp_path):
    eng = E.TFKerasModel(_cfg('unet'))
    d = tmp_path / 'checkpoints'
    d.mkdir()
    for f in ('ckpt-300.index', 'ckpt-300.data-00000-of-00001', 'ckpt-100.index', 'ckpt-100.data-00000-of-00001', 'ckpt-200.npz',
              'checkpoint', 'ckpt-abc.index', 'other-5.index'):
        (d / f).write_bytes(b'')
    got = eng.get_ckpts(str(d))
    assert list(got) == [100, 200, 300] and got[100] == str(d / 'ckpt-100') and got[200] == str(d / 'ckpt-200')
    assert eng.list_ckpts(str(d)) == got and eng.get_ckpts(str(tmp_path / 'nope')) == {}
    assert E.solve_learning_rate_scheduler(None) is None
    with pytest.raises(ValueError):
        E.solve_learning_rate_scheduler('3')


def test_input_shape_is_peeked_only_from_reiterable_datasets():
    x = np.zeros((2, 32, 32, 3), np.float32)
    assert E.TFKerasModel._peek_input_shape([(x, x[..., 0])]) == (None, 32, 32, 3)
    gen = ((x, x[..., 0]) for _ in range(3))
    with pytest.raises(ValueError, match='input_shape'):
        E.TFKerasModel._peek_input_shape(gen)
    assert len(list(gen)) == 3                      # nothing was consumed
