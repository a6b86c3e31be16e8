"""CPU-side checks: the C-ABI library loads and exports every symbol include/dnnca.h
declares, and the host-side lowering (model classes -> static plan) has the structure of
the reference graphs.  No kernels are launched here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def libpath():
    from dnncancerannotator_b200 import build
    return build.build()          # nvcc cross-compiles sm_100a without a GPU


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'dnnca.h')).read()
    return sorted(set(re.findall(r'DNNCA_API\s+[\w\s\*]+?\b(dnnca_\w+)\s*\(', src)))


def test_library_exports_every_declared_symbol(libpath):
    declared = header_symbols()
    assert len(declared) >= 27
    out = subprocess.run(['nm', '-D', '--defined-only', libpath], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r' T (dnnca_\w+)', out))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    # nothing but the C ABI leaks out of the library
    assert all(s.startswith('dnnca_') for s in re.findall(r' T (\w+)', out))


def test_ctypes_binding_covers_the_header(libpath):
    from dnncancerannotator_b200 import native
    assert sorted(native.exported_symbols()) == header_symbols()
    lib = native.lib()
    assert lib.dnnca_version() == 100
    assert ctypes.sizeof(native.Tensor) == 40 and ctypes.sizeof(native.LabelStats) == 16
    assert ctypes.sizeof(native.LossConfig) == 20


def test_sass_is_sm100a(libpath):
    out = subprocess.run(['cuobjdump', '-lelf', libpath], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from dnncancerannotator_b200 import native
    monkeypatch.setattr(native, 'LIB_PATH', str(tmp_path / 'nope.so'))
    monkeypatch.setattr(native, '_lib', None)
    with pytest.raises(native.DnncaError, match='no CPU or PyTorch fallback'):
        native.lib()


def test_no_cuda_device_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip('needs a CPU-only host')
    from dnncancerannotator_b200 import native
    from dnncancerannotator_b200.models import tf_models
    m = tf_models.UNetAnnotator(3, 2, 2, 3, 1, padding='same')
    with pytest.raises(native.DnncaError, match='no CPU fallback'):
        m(np.zeros((1, 16, 16, 3), np.float32))


def test_product_never_imports_the_oracle():
    """Only tests/ (incl. tests/tools), __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/: neither the package
    nor the scripts under tools/ import it."""
    bad = []
    for top in ('dnncancerannotator_b200', 'tools'):
        for d, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith('.py') and re.search(r'^\s*(from|import)\s+oracle\b', open(os.path.join(d, f)).read(), re.M):
                    bad.append(os.path.join(top, f))
    assert not bad, bad
    # bench.py: the oracle appears only inside the CPU comparator / reference arm
    src = open(os.path.join(ROOT, 'bench.py')).read()
    for m in re.finditer(r'^\s*(from|import)\s+oracle\b.*$', src, re.M):
        head = src[:m.start()]
        fn = re.findall(r'^def (\w+)\(', head, re.M)[-1]
        assert fn in ('cpu_reference_rate', 'run_reference'), fn


# ---- config surface -------------------------------------------------------------------
def test_layered_config_loading_matches_reference_semantics():
    from dnncancerannotator_b200.utils.load import load_config
    c = os.path.join(ROOT, 'configs')
    cfg = load_config([f'{c}/unet.yaml', f'{c}/additionals/data_options.yaml', f'{c}/additionals/deploy_options.yaml',
                       f'{c}/additionals/multigpu.yaml', f'{c}/additionals/train_batch28.yaml',
                       f'{c}/additionals/leakyReLU.yaml', f'{c}/additionals/kernel_regularizer.yaml'])
    assert cfg['model'] == 'UNetAnnotator' and cfg['model_options']['n_filters_first'] == 3
    assert cfg['deploy_options']['enable_multigpu'] is True              # dotted-key overlay (load.py:44-57)
    assert cfg['data_options']['train'] == {'batch_size': 28}           # whole-value overwrite like the reference
    assert cfg['model_options']['activation']['config']['alpha'] == 0.3
    assert cfg['deploy_options']['loss']['config']['weight_mul'] == 3.0
    assert cfg['data_options']['eval']['batch_size'] == 64


def test_loss_registry_and_config_keys():
    from dnncancerannotator_b200.utils import losses
    l = losses.get({'class_name': 'WeightedCrossentropy', 'config': {'weight_mul': 3.0}})
    assert l.get_config() == dict(weight=None, weight_add=0.0, weight_mul=3.0, label_smoothing=False,
                                  label_smoothing_filter_size=6, label_smoothing_sigma=3)
    cfg = l.native_config(1000)
    assert cfg.has_weight == 0 and cfg.weight_mul == 3.0 and abs(cfg.grad_scale - 1e-3) < 1e-9
    assert losses.get('WeightedCrossentropy').weight_mul == 1.0
    ls = losses.WeightedCrossentropy(label_smoothing=True)         # configs/additionals/enable_label_smoothing.yaml
    assert ls.get_config()['label_smoothing'] is True and ls.label_smoothing_filter_size == 6
    with pytest.raises(ValueError):                                 # the smoothing runs on the device only
        ls.prepare_labels(np.zeros((1, 8, 8), np.float32))
    with pytest.raises(ValueError):
        losses.get('mse')


# ---- lowering --------------------------------------------------------------------------
def cpu_plan(model, B, H, C, dtype=torch.bfloat16, training=True):
    from dnncancerannotator_b200 import runtime as R
    model.build((None, H, H, C))
    model.params.materialize(torch.device('cpu'))
    plan = R.Plan(model.params, B, H, H, C, dtype, torch.device('cpu'))
    model._emit(plan)
    plan.allocate(training)
    return plan


def op_counts(plan):
    out = {}
    for op in plan.ops:
        out[type(op).__name__] = out.get(type(op).__name__, 0) + 1
    return out


def test_unet_yaml_lowering_structure():
    from dnncancerannotator_b200.models import tf_models
    from dnncancerannotator_b200 import runtime as R
    from dnncancerannotator_b200.utils.load import load_config
    cfg = load_config(os.path.join(ROOT, 'configs', 'unet.yaml'))
    m = getattr(tf_models, cfg['model'])(**cfg['model_options'])
    plan = cpu_plan(m, 2, 64, 3)
    assert m.count_params() == 8740
    assert op_counts(plan) == {'ConvertOp': 1, 'ConvOp': 12, 'PoolOp': 3, 'TConvOp': 3}
    convs = [op for op in plan.ops if isinstance(op, R.ConvOp)]
    assert [(o.x.c, o.x2.c if o.x2 else 0, o.y.c) for o in convs] == [
        (3, 0, 3), (3, 0, 3), (3, 0, 6), (6, 0, 6), (6, 0, 12), (12, 0, 12),
        (12, 12, 12), (12, 0, 12), (6, 6, 6), (6, 0, 6), (3, 3, 3), (3, 0, 3)]
    assert not convs[0].x.needs_grad                          # input gets no gradient (first dgrad skipped)
    # tf.concat([tconv, skip]) is virtual: the first decoder conv of a level reads the transposed-conv
    # output (x) and the encoder's skip tensor (x2); the skip tensor is also the max-pool's input
    pools = [op for op in plan.ops if isinstance(op, R.PoolOp)]
    tconvs = [op for op in plan.ops if isinstance(op, R.TConvOp)]
    dec0 = [o for o in convs if o.x2 is not None]
    for pool, tconv, conv in zip(reversed(pools), tconvs, dec0):
        assert pool.x.skip_consumed and conv.x2 is pool.x and conv.x is tconv.y
        assert all(t.coff == 0 and t.buf.c == t.c for t in (conv.x, conv.x2, conv.y))   # every tensor is dense
    assert plan.features.shape == (2, 64, 64, 3) and plan.features.act is not None
    assert plan.head == ('head/kernel', 'head/bias')


def test_unet_big_and_mulmo_lowering_structure():
    from dnncancerannotator_b200.models import tf_models
    from dnncancerannotator_b200 import runtime as R
    from dnncancerannotator_b200.utils.load import load_config
    cfg = load_config(os.path.join(ROOT, 'configs', 'unet_big.yaml'))
    m = getattr(tf_models, cfg['model'])(**cfg['model_options'])
    plan = cpu_plan(m, 1, 32, 3, training=False)
    assert m.count_params() == 15848385 and m.count_params(True) == 15836865
    assert op_counts(plan) == {'ConvertOp': 1, 'ConvOp': 16, 'BNOp': 24, 'PoolOp': 4, 'TConvOp': 4}
    cfg = load_config(os.path.join(ROOT, 'configs', 'mulmo_unet.yaml'))
    m = getattr(tf_models, cfg['model'])(**cfg['model_options'])
    plan = cpu_plan(m, 1, 32, 3, training=False)
    assert m.count_params() == 1719089 and m.count_params(True) == 1713329
    # bf16: one dense single-channel input buffer per modality so the 1->16 first convs run on the tensor cores
    assert op_counts(plan) == {'ConvertOp': 3, 'ConvOp': 32, 'BNOp': 48, 'PoolOp': 12, 'TConvOp': 4}
    first = [op for op in plan.ops if isinstance(op, R.ConvOp)][0]
    assert first.x.c == 1 and first.x.buf.c == 1                 # inputs[..., m:m+1] as a dense tensor (row-Toeplitz kernels)
    bott = [b for b in plan.bufs if b.name == 'bottleneck'][0]
    assert (bott.h, bott.w, bott.c) == (2, 2, 384)               # concat of the three 128-channel encoders


def test_multiresunet_lowering_structure():
    from dnncancerannotator_b200.models import tf_models
    m = tf_models.MultiResUnet(None, None, 5)
    plan = cpu_plan(m, 1, 32, 5, training=False)
    assert m.count_params() == 7262996
    c = op_counts(plan)
    assert c['_FoldedConv'] == 56 and c['_FoldedTConv'] == 4 and c['PoolOp'] == 4 and c['AddReluAffineOp'] == 19
    # odd block widths live in buffers whose concat segments are padded to multiples of 8 channels
    # (51 = 8|17|26 -> 8|24|32 = 64, ..., 853 -> 864) so that every conv can take the tensor-core path
    from dnncancerannotator_b200.models.tf_models.multiresunet import _FoldedConv, _FoldedTConv
    convs = [op for op in plan.ops if isinstance(op, _FoldedConv)]
    assert all(op.x.c % 8 == 0 and op.y.c % 8 == 0 and op.y.coff % 8 == 0 for op in convs)
    widths = sorted({b.c for b in plan.bufs if b.name == 'mres_out'})
    assert widths == [64, 120, 224, 432, 864]
    assert [op.xs.c for op in plan.ops if isinstance(op, _FoldedTConv)] == [853, 426, 212, 105]
    m0 = convs[1].xs.chmap()                                     # the network input: 5 modalities in 8 physical channels
    assert list(m0) == [0, 1, 2, 3, 4, -1, -1, -1]
    cat = [op for op in plan.ops if type(op).__name__ == 'AddReluAffineOp'][0]
    assert cat.b.c == 64 and cat.a.c == 64


def test_multiresblock_is_exported_and_usable():
    """tf_models/__init__.py:2 exports MultiResBlock(U, inp, alpha) as a graph-building helper."""
    from dnncancerannotator_b200.models import tf_models
    from dnncancerannotator_b200.models.tf_models import multiresunet as mr
    m = tf_models.MultiResUnet(None, None, 5)
    b = mr._Builder(m)
    out = tf_models.MultiResBlock(32, mr.Sym(b, None, 5))
    assert out.c == 51 and m.params.count() > 0
    with pytest.raises(TypeError):
        tf_models.MultiResBlock(32, object())


def test_unsupported_geometry_is_rejected():
    from dnncancerannotator_b200.models import tf_models
    with pytest.raises(NotImplementedError, match="padding='valid'"):
        tf_models.UNetAnnotator(3, 2, 2, 3, 1)                   # reference default padding
    with pytest.raises(NotImplementedError):
        tf_models.UNetAnnotator(3, 2, 3, 3, 1, padding='same')   # rate 3


def test_weights_roundtrip_and_checkpoint_naming(tmp_path):
    from dnncancerannotator_b200.models import tf_models
    m = tf_models.UNetAnnotator(4, 2, 2, 3, 1, bn=True, padding='same')
    m.build((None, 32, 32, 3))
    w = m.get_weights()
    w['head/bias'] = np.array([0.5], np.float32)
    m.set_weights(w)
    os.makedirs(tmp_path / 'checkpoints')
    m.save_weights(str(tmp_path / 'checkpoints' / 'ckpt-5000'))
    m.save_weights(str(tmp_path / 'checkpoints' / 'ckpt-10000'))
    assert list(m.list_checkpoints(str(tmp_path))) == [5000, 10000]
    m2 = tf_models.UNetAnnotator(4, 2, 2, 3, 1, bn=True, padding='same', seed=9)
    m2.build((None, 32, 32, 3))
    m2.load_weights(m.list_checkpoints(str(tmp_path))[10000])
    for k, v in m.get_weights().items():
        np.testing.assert_array_equal(v, m2.get_weights()[k])
    with pytest.raises(KeyError):
        m2.set_weights({'nope': np.zeros(1)})
    with pytest.raises(ValueError):
        m2.set_weights({'head/bias': np.zeros(2)})


def test_bind_host_to_gpu_is_a_noop_without_nvml_or_gpu():
    """parallel.bind_host_to_gpu must never raise or change the affinity when there is no GPU / NVML to ask."""
    import os
    from dnncancerannotator_b200.parallel import bind_host_to_gpu
    before = os.sched_getaffinity(0)
    assert bind_host_to_gpu(0) is None
    assert os.sched_getaffinity(0) == before


# ---- round 2: boundary completions ---------------------------------------------------------
def test_load_status_and_model_save(tmp_path):
    """engine.py:75 ``load_weights(...).assert_existing_objects_matched()`` and engine.py:226 ``model.save``."""
    import json
    from dnncancerannotator_b200.models import tf_models
    m = tf_models.UNetAnnotator(4, 2, 2, 3, 1, bn=True, padding='same')
    m.build((None, 32, 32, 3))
    m.compile(loss={'class_name': 'WeightedCrossentropy', 'config': {'weight_mul': 3.0}})
    d = m.save(str(tmp_path / 'saved'))
    cfg = json.load(open(os.path.join(d, 'config.json')))
    assert cfg['class_name'] == 'UNetAnnotator' and cfg['config']['n_filters_first'] == 4
    assert cfg['loss']['config']['weight_mul'] == 3.0 and cfg['input_shape'] == [None, 32, 32, 3]
    m2 = getattr(tf_models, cfg['class_name'])(**cfg['config'], seed=5)
    m2.build(tuple(cfg['input_shape']))
    st = m2.load_weights(d)                                      # the directory written by save()
    assert st.assert_existing_objects_matched() is st and st.assert_consumed() is st
    for k, v in m.get_weights().items():
        np.testing.assert_array_equal(v, m2.get_weights()[k])
    # a checkpoint that lacks variables of the model / holds foreign ones
    w = m.get_weights()
    del w['head/bias']
    np.savez(str(tmp_path / 'partial.npz'), **w, extra=np.zeros(3, np.float32))
    st = m2.load_weights(str(tmp_path / 'partial'))
    assert st.missing == ['head/bias'] and st.unused == ['extra']
    with pytest.raises(AssertionError):
        st.assert_existing_objects_matched()
    assert st.expect_partial() is st


def test_unknown_layer_kwargs_and_frozen_layers_are_rejected():
    from dnncancerannotator_b200.models.tf_models import components
    with pytest.raises(TypeError, match='unexpected keyword'):
        components.Downsample(4, 2, 3, 1, False, padding='same', dilation=2)
    with pytest.raises(NotImplementedError, match='trainable=False'):
        components.Downsample(4, 2, 3, 1, False, padding='same', trainable=False)


def test_ready_frontier_and_bucket_schedule():
    """Data-parallel overlap schedule: a bucket of the flat gradient buffer is issued once every gradient in it has
    been written; the frontier falls monotonically through the backward pass and ends at 0."""
    from dnncancerannotator_b200.models import tf_models
    from dnncancerannotator_b200.parallel import GradAllReduce, bucket_ranges
    from dnncancerannotator_b200.utils.load import load_config
    for cfgname in ('unet', 'unet_big', 'mulmo_unet'):
        cfg = load_config(os.path.join(ROOT, 'configs', cfgname + '.yaml'))
        m = getattr(tf_models, cfg['model'])(**cfg['model_options'])
        plan = cpu_plan(m, 1, 32, 3)
        ready = plan.ready_frontier()
        ps = m.params
        assert len(ready) == len(plan.ops)
        assert all(a >= b for a, b in zip(ready, ready[1:])) and ready[-1] == 0
        # before the backward pass only the head's gradients (the highest offsets) are complete
        head_off = ps.specs['head/kernel']['offset']
        assert plan.pending_before_backward <= head_off
        # replay the schedule with a recording stand-in for the collective
        dp = GradAllReduce.__new__(GradAllReduce)
        dp.world_size, dp.bucket_elems, dp.min_buckets, dp.group = 2, (8 << 20) // 4, 4, None
        issued = []

        class W:
            def wait(self):
                pass
        import torch.distributed as dist
        orig = dist.all_reduce
        dist.all_reduce = lambda t, op=None, group=None, async_op=False: (issued.append(t.numel()), W())[1]
        try:
            flat = ps.grads_full
            dp.begin(flat)
            written = set()
            dp.launch_ready(plan.pending_before_backward)
            for i, op in enumerate(reversed(plan.ops)):
                written.update(op.grad_params())
                dp.launch_ready(ready[i])
                # every gradient inside an issued bucket has been written (or is the head's / the loss slot)
                for _, (a, b) in dp.launch_log:
                    for n, sp in ps.specs.items():
                        if sp['trainable'] and a <= sp['offset'] < b and not n.startswith('head/'):
                            assert n in written, (cfgname, n, a, b)
            early = len(dp.launch_log)
            dp.finish()
        finally:
            dist.all_reduce = orig
        assert sum(issued) == flat.numel() and len(issued) >= 4
        assert early >= len(issued) - 1, (cfgname, early, len(issued))      # only the last bucket waits for the end
        covered = sorted(i for _, (a, b) in dp.launch_log for i in (a, b))
        assert covered[0] == 0 and covered[-1] == flat.numel()


def test_multiresunet_training_plan_structure_and_variable_maps():
    """The channel-padded training plan of MultiResUnet (multires_train.py), inspected on the CPU: it declares exactly the
    variables of the model, the index maps place every logical element at one physical position (holes read 0), the
    gradient gather is the inverse, and the op list has the reference's 56 conv2d_bn + 28 full BatchNorms + 4 ConvT."""
    import numpy as np
    import torch
    from collections import Counter
    from dnncancerannotator_b200.models import tf_models
    from dnncancerannotator_b200.models.tf_models import multires_train as T
    m = tf_models.MultiResUnet(None, None, 5)
    m.build((None, 32, 32, 5))
    cpu = torch.device('cpu')
    m.params.materialize(cpu)
    plan = T.emit_training_plan(m, 2, 32, 32, device=cpu)
    lg, ph, mp = plan.logical, plan.phys, plan.maps
    assert set(plan.links) == set(lg.specs)
    lg.params.copy_(torch.randn_like(lg.params))
    lg.state.copy_(torch.rand_like(lg.state) + 0.5)

    def gather(src, idx):               # what dnnca_gather_f32 computes
        out = torch.zeros(idx.numel())
        ok = idx >= 0
        out[ok] = src[idx[ok].long()]
        return out
    ph.params.copy_(gather(lg.params, mp['p2l_t']))
    ph.state.copy_(gather(lg.state, mp['p2l_s']))
    for name, index in plan.links.items():
        a, b = lg.view(name).numpy(), ph.view(name).numpy()
        sel = b[np.ix_(*index)] if index is not None else b
        assert np.array_equal(sel, a), name
        assert np.count_nonzero(b) == np.count_nonzero(a), name          # holes are zeros
        assert all(d % 16 == 0 or d in (1, 2, 3) for d in b.shape if d > 3), (name, b.shape)
    for src, dst, idx in ((ph.params, lg.params, mp['l2p_t']), (ph.state, lg.state, mp['l2p_s'])):
        back, valid = gather(src, idx), idx >= 0
        assert torch.equal(back[valid], dst[valid])
    assert int((mp['l2p_t'] >= 0).sum()) == lg.count(True) and int((mp['l2p_s'] >= 0).sum()) == lg.count(False)
    kinds = Counter(type(o).__name__ for o in plan.ops)
    assert kinds['ConvOp'] == 56 and kinds['BNActOp'] == 56 and kinds['BNOp'] == 28 and kinds['TConvOp'] == 4 and kinds['PoolOp'] == 4
    assert plan.ready_frontier() == [lg.grads_full.numel()] * len(plan.ops)     # no gradient bucket leaves before the gather


def test_save_and_load_model_without_a_device(tmp_path):
    """engine.py:226 ``model.save(path)`` + ``load_model``: class, constructor config, loss and optimizer settings and every
    variable come back from the directory alone (host side; the device round trip incl. Adam slots is a -m gpu test)."""
    import json
    import numpy as np
    from dnncancerannotator_b200.keras_like import load_model
    from dnncancerannotator_b200.models import tf_models
    cases = [('UNetAnnotator', dict(n_filters_first=3, n_downsample=2, rate=2, kernel_size=3, conv_stride=1, padding='same',
                                   bn=True, activation=dict(class_name='LeakyReLU', config=dict(alpha=0.3)))),
             ('MulmoUNetAnnotator', dict(n_filters_first=4, n_downsample=2, rate=2, kernel_size=3, conv_stride=1, padding='same')),
             ('MultiResUnet', dict(height=None, width=None, n_channels=5))]
    for i, (name, opts) in enumerate(cases):
        m = getattr(tf_models, name)(**opts, dtype='fp32', seed=3)
        m.build((None, 32, 32, 5 if name == 'MultiResUnet' else 3))
        m.compile(optimizer=dict(learning_rate=5e-4), loss=dict(class_name='WeightedCrossentropy', config=dict(weight_mul=3.0)))
        d = m.save(str(tmp_path / f'saved{i}'))
        cfg = json.load(open(os.path.join(d, 'config.json')))
        assert cfg['class_name'] == name and cfg['compute_dtype'] == 'fp32'
        m2 = load_model(d)
        assert type(m2) is type(m) and m2.get_config() == m.get_config()
        assert m2.optimizer['learning_rate'] == 5e-4 and m2.loss.weight_mul == 3.0
        w1, w2 = m.get_weights(), m2.get_weights()
        assert list(w1) == list(w2) and all(np.array_equal(w1[k], w2[k]) for k in w1)
    with pytest.raises(ValueError):
        json.dump(dict(class_name='NoSuchModel', config={}), open(os.path.join(d, 'config.json'), 'w'))
        load_model(d)


def test_integration_doc_names_every_exported_symbol():
    """INTEGRATION.md lists the reference call site of every entry point include/dnnca.h declares (names may be grouped as
    ``dnnca_x_{a,b}`` or ``dnnca_x_*``)."""
    import itertools
    from dnncancerannotator_b200 import native
    integ = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    named, wild = set(), []
    for tok in re.findall(r'`(dnnca_[A-Za-z0-9_{},*]+)`', integ):
        parts = re.split(r'(\{[^}]*\})', tok)
        options = [p[1:-1].split(',') if p.startswith('{') else [p] for p in parts]
        for combo in itertools.product(*options):
            name = ''.join(combo)
            if name.endswith('*'):
                wild.append(name[:-1])
            else:
                named.add(name)
    missing = [s for s in native.exported_symbols() if s not in named and not any(s.startswith(w) for w in wild)]
    assert not missing, missing
    unknown = sorted(n for n in named if n not in native.exported_symbols() and not n.endswith('_t'))       # (types)
    assert not unknown, unknown                       # the document names nothing the library does not export


def test_pack_labels_host_side():
    """data_tail.pack_labels: bit order = numpy.packbits (what dnnca_unpack_label_bits expands), uint8 {0,255} and float
    {0,1} masks give the same bits, anything fractional / ragged is refused."""
    from hypothesis import given, settings, strategies as st
    from dnncancerannotator_b200 import data_tail

    @settings(max_examples=25, deadline=None)
    @given(st.integers(1, 3), st.integers(1, 9), st.integers(1, 5), st.integers(0, 2 ** 31 - 1))
    def roundtrip(b, h, w8, seed):
        rng = np.random.default_rng(seed)
        y = (rng.random((b, h, 8 * w8)) > 0.7)
        for arr in (y.astype(np.float32), (y * 255).astype(np.uint8), torch.from_numpy(y.astype(np.float32))):
            p = data_tail.pack_labels(arr, pinned=False)
            assert p.shape == y.shape and p.nbytes * 8 == y.size and p.bits.dtype == torch.uint8
            assert np.array_equal(np.unpackbits(p.bits.numpy(), axis=-1).astype(bool), y)
    roundtrip()
    y = np.zeros((1, 4, 16), np.float32)
    y[0, 0, 0] = 0.5
    with pytest.raises(ValueError):
        data_tail.pack_labels(y, pinned=False)
    with pytest.raises(ValueError):
        data_tail.pack_labels(np.zeros((1, 4, 12), np.float32), pinned=False)
    with pytest.raises(ValueError):
        data_tail.pack_labels(np.zeros((4, 16), np.float32), pinned=False)
