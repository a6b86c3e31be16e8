"""Data parallelism of the CUDA path on real GPUs (engine.py:260-263): two ranks over NCCL on one box.

Parity statement: an R-rank run with per-rank batches must equal the oracle evaluated on R independent sub-batches
(rank-local BatchNorm statistics, rank-local positive-rate loss weight) whose gradients are averaged; variables are
mirrored from rank 0; the reported loss is the global mean.  Run with ``gpurun --gpus 2 -- python -m pytest
tests/test_dp_nccl.py -m gpu`` (skipped on a single-GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

OPTS = dict(n_filters_first=4, n_downsample=2, rate=2, kernel_size=3, conv_stride=1, bn=True, padding='same')
LOSS = dict(weight_mul=3.0)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, graph, p2p, out):
    import faulthandler
    import sys
    import torch.distributed as dist
    logdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    os.makedirs(logdir, exist_ok=True)
    log = open(os.path.join(logdir, f'dp_test_{mode}_{int(graph)}{int(p2p)}_rank{rank}.log'), 'w')
    faulthandler.dump_traceback_later(90, file=log, exit=True)      # a hung collective must not eat the GPU budget

    def say(msg):
        log.write(msg + '\n')
        log.flush()
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    os.environ['DNNCA_DP_GRAPH'] = '1' if graph else '0'
    os.environ['DNNCA_P2P'] = '1' if p2p else '0'      # gradient exchange: NVLink peer memory inside Adam / bucketed NCCL
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from dnncancerannotator_b200.models import tf_models
        from dnncancerannotator_b200.synthetic import make_slices
        from oracle import ref_models as rm
        from oracle import ref_ops as ops
        H = 32
        m = tf_models.UNetAnnotator(**OPTS, dtype=mode, seed=rank)          # replicas start DIFFERENT ...
        m.build((None, H, H, 3))
        m.compile(loss=dict(class_name='WeightedCrossentropy', config=LOSS))
        ref0 = rm.build_model('UNetAnnotator', OPTS, (None, H, H, 3), seed=10 + rank)
        ref0.randomize_bn(seed=3 + rank)
        m.set_weights(ref0.get_weights())
        say('model built')
        m.enable_data_parallel(bucket_bytes=1024)                            # ... and are mirrored from rank 0; many buckets
        say('dp enabled')
        w_after_sync = m.get_weights()
        batches = [make_slices(3, H, H, 3, seed=1234 + r) for r in range(world)]
        x, y = batches[rank]
        # oracle: every replica's sub-batch on its own (own BN statistics, own loss weight), gradients averaged
        ref = rm.build_model('UNetAnnotator', OPTS, (None, H, H, 3), seed=10)
        ref.randomize_bn(seed=3)
        per = [ref.train_step_grads(bx, by, LOSS) for bx, by in batches]
        avg = {k: sum(p['grads'][k] for p in per) / world for k in ref.trainable}
        mean_loss = float(np.mean([p['loss'] for p in per]))
        say('oracle done')
        losses = []
        for i in range(4):                                                   # eager, eager, capture + replay, replay
            losses.append(float(m.train_step(x, y)))
            say(f'step {i} loss {losses[-1]}')
        g = m.get_grads()                                                    # all-reduced gradients of the LAST step
        w = m.get_weights()
        # oracle weights after ONE Adam step with the averaged gradient
        w1 = {}
        for k in ref.trainable:
            w1[k], _, _ = ops.adam_step(ref.weights[k], avg[k], torch.zeros_like(avg[k]), torch.zeros_like(avg[k]), 1)
        res = dict(w_sync=w_after_sync, losses=losses, mean_loss=mean_loss,
                         log=[(f, b) for f, b in m._dp.launch_log], grads=g, w=w,
                         avg={k: v.numpy() for k, v in avg.items()}, w1={k: v.numpy() for k, v in w1.items()},
                         ref_w0={k: v.numpy() for k, v in ref.weights.items()})
        # first-step weights: re-run one step from the synced weights
        m2 = tf_models.UNetAnnotator(**OPTS, dtype=mode, seed=0)
        m2.build((None, H, H, 3))
        m2.compile(loss=dict(class_name='WeightedCrossentropy', config=LOSS))
        m2.set_weights(ref.get_weights())
        m2.enable_data_parallel(bucket_bytes=1024)
        say('m2 ready')
        l1 = float(m2.train_step(x, y))
        say('m2 stepped')
        res.update(l1=l1, g1=m2.get_grads(), w_step1=m2.get_weights())
        out[rank] = res
        say('results stored')
        assert (m._p2p is not None) == bool(p2p)
        m.close()                              # graphs holding captured NCCL collectives / peer mappings go before the communicator
        m2.close()
        say('graphs released')
    finally:
        dist.destroy_process_group()
        say('group destroyed')
        faulthandler.cancel_dump_traceback_later()


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64).ravel() - np.asarray(b, np.float64).ravel()) /
                 max(np.linalg.norm(np.asarray(b, np.float64).ravel()), 1e-30))


@pytest.mark.parametrize('mode,graph,p2p', [('fp32', True, False), ('fp32', False, False), ('fp32', True, True), ('fp32', False, True),
                                            ('bf16', True, True), ('bf16', True, False)])
def test_two_rank_step_matches_oracle_subbatch_average(mode, graph, p2p):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (gpurun --gpus 2)')
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), mode, graph, p2p, out), nprocs=world, join=True)
    a, b = out[0], out[1]
    tol = 2e-3 if mode == 'fp32' else 0.3
    # mirrored variables: rank 1 took rank 0's values when data parallelism was enabled
    for k in a['w_sync']:
        np.testing.assert_array_equal(a['w_sync'][k], b['w_sync'][k])
        np.testing.assert_array_equal(a['w_sync'][k], a['ref_w0'][k])
    # one step from the oracle's weights: all-reduced gradients == average of the oracle's per-replica gradients
    names = [k for k in a['avg'] if not k.endswith('/tconv/bias')]
    got = np.concatenate([a['g1'][k].ravel() for k in names])
    want = np.concatenate([a['avg'][k].ravel() for k in names])
    assert _rel(got, want) <= tol, _rel(got, want)
    for k in names:
        np.testing.assert_array_equal(a['g1'][k], b['g1'][k])              # both ranks hold the same reduced gradient
    assert abs(a['l1'] - a['mean_loss']) <= (1e-4 if mode == 'fp32' else 2e-2) * abs(a['mean_loss'])
    assert a['l1'] == b['l1']                                              # the loss scalar travels in the all-reduce
    if mode == 'fp32':
        for k in names:
            assert _rel(a['w_step1'][k], a['w1'][k]) < 5e-3 or np.abs(a['w_step1'][k] - a['w1'][k]).max() < 2e-4, k
    # replicas stay mirrored over the four steps (eager, eager, captured graph, replay)
    for k in a['w']:
        np.testing.assert_array_equal(a['w'][k], b['w'][k]) if 'moving' not in k else None
    assert all(np.isfinite(a['losses'])) and a['losses'] == b['losses']
    # overlap schedule: several buckets, all but the last issued before the backward pass ended (frontier > 0)
    log = a['log']
    if p2p:
        assert log == []                       # no NCCL collective on the step path at all
        return
    # (with these 256-element test buckets the first layer's kernel + bias spill into the last two)
    assert len(log) >= 4 and sum(1 for f, _ in log if f > 0) >= len(log) - 2, log


def _abi_worker(rank, world, idfile, out):
    """C-ABI exchange (include/dnnca.h dnnca_nccl_*) without torch.distributed: the unique id travels through a file."""
    import ctypes as C
    import faulthandler
    import sys
    import time
    faulthandler.dump_traceback_later(90, exit=True)
    from dnncancerannotator_b200 import native as N
    torch.cuda.set_device(rank)
    lib = N.lib()
    if rank == 0:
        buf = C.create_string_buffer(128)
        N.check(lib.dnnca_nccl_unique_id(buf), 'nccl_unique_id')
        with open(idfile + '.tmp', 'wb') as f:
            f.write(buf.raw)
        os.replace(idfile + '.tmp', idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            assert time.time() - t0 < 60
            time.sleep(0.05)
    uid = open(idfile, 'rb').read()
    assert len(uid) == 128
    comm = C.c_void_p()
    N.check(lib.dnnca_nccl_comm_init_rank(C.byref(comm), world, uid, rank), 'nccl_comm_init_rank')
    rng = np.random.default_rng(100 + rank)
    g = rng.normal(size=10007).astype(np.float32)
    gd = torch.from_numpy(g).cuda()
    w = torch.full((333,), float(rank + 1), device='cuda')
    # mirrored variables from rank 0, then the flat gradient buffer in reverse-order buckets on the current stream
    N.check(lib.dnnca_nccl_broadcast(comm, N.stream_ptr(), N.ptr(w), w.numel() * 4, 0), 'nccl_broadcast')
    from dnncancerannotator_b200.parallel import bucket_ranges
    for a, b in bucket_ranges(gd.numel(), 4096):
        N.check(lib.dnnca_nccl_allreduce_bucket(comm, N.stream_ptr(), C.c_void_p(gd.data_ptr() + 4 * a), b - a, N.F32),
                'nccl_allreduce_bucket')
    hb = torch.from_numpy(g).cuda().bfloat16()
    N.check(lib.dnnca_nccl_allreduce_bucket(comm, N.stream_ptr(), N.ptr(hb), hb.numel(), N.BF16), 'nccl_allreduce_bucket')
    torch.cuda.synchronize()
    rc = lib.dnnca_nccl_allreduce_bucket(comm, N.stream_ptr(), None, 4, N.F32)          # bad argument: loud, not a hang
    out[rank] = dict(g=g, reduced=gd.cpu().numpy(), w=w.cpu().numpy(), bf16=hb.float().cpu().numpy(), bad_rc=int(rc))
    N.check(lib.dnnca_nccl_comm_destroy(comm), 'nccl_comm_destroy')
    faulthandler.cancel_dump_traceback_later()


def test_c_abi_nccl_wrappers_two_ranks(tmp_path):
    """include/dnnca.h `dnnca_nccl_*`: unique id -> communicator -> broadcast + bucketed SUM all-reduce"""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (gpurun --gpus 2)')
    import torch.multiprocessing as mp
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_abi_worker, args=(world, str(tmp_path / 'nccl_id'), out), nprocs=world, join=True)
    a, b = out[0], out[1]
    want = a['g'] + b['g']
    np.testing.assert_array_equal(a['reduced'], b['reduced'])
    np.testing.assert_allclose(a['reduced'], want, rtol=0, atol=1e-6)
    np.testing.assert_array_equal(a['w'], np.ones(333, np.float32))
    np.testing.assert_array_equal(b['w'], np.ones(333, np.float32))
    np.testing.assert_array_equal(a['bf16'], b['bf16'])
    np.testing.assert_allclose(a['bf16'], want, rtol=2e-2, atol=2e-2)
    assert a['bad_rc'] == -1 and b['bad_rc'] == -1


def _worker_multires(rank, world, port, out):
    """MultiResUnet under data parallelism: its training plan computes on channel-padded copies of the variables, so no
    gradient bucket may leave before the gradients were gathered back (``MultiResTrainPlan.ready_frontier``)."""
    import faulthandler
    import torch.distributed as dist
    logdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    os.makedirs(logdir, exist_ok=True)
    log = open(os.path.join(logdir, f'dp_test_multires_rank{rank}.log'), 'w')
    faulthandler.dump_traceback_later(120, file=log, exit=True)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from dnncancerannotator_b200.models import tf_models
        from dnncancerannotator_b200.synthetic import make_slices
        from oracle import ref_models as rm
        H = 32
        mo = dict(height=None, width=None, n_channels=5)
        ref = rm.build_model('MultiResUnet', mo, None, seed=0)
        ref.randomize_bn(seed=1)
        m = tf_models.MultiResUnet(**mo, dtype='fp32', seed=rank)
        m.build((None, H, H, 5))
        m.compile(loss=dict(class_name='WeightedCrossentropy', config=LOSS))
        if rank == 0:
            m.set_weights(ref.get_weights())
        m.enable_data_parallel(bucket_bytes=1 << 20)                         # variables mirrored from rank 0
        batches = [make_slices(2, H, H, 5, seed=(1235, 35)[r]) for r in range(world)]
        x, y = batches[rank]
        per = [ref.train_step_grads(bx, by, LOSS) for bx, by in batches]
        avg = {k: (sum(p['grads'][k] for p in per) / world).numpy() for k in ref.trainable}
        # the same kernels WITHOUT data parallelism on this rank's sub-batch: their cross-rank mean is what the all-reduce
        # must deliver to the last bit of fp32 summation (61 BatchNorm layers deep a max-pool / relu near-tie decided
        # differently by two implementations moves the gradient by 1e-2, so the oracle bound below is loose)
        solo = tf_models.MultiResUnet(**mo, dtype='fp32', seed=0)
        solo.build((None, H, H, 5))
        solo.compile(loss=dict(class_name='WeightedCrossentropy', config=LOSS))
        solo.set_weights(ref.get_weights())
        solo.forward_backward(x, y)
        flat = solo.params.grads.clone()
        dist.all_reduce(flat)
        flat /= world
        solo_avg = {k: flat[sp['offset']:sp['offset'] + sp['numel']].view(sp['shape']).cpu().numpy()
                    for k, sp in solo.params.specs.items() if sp['trainable']}
        l1 = float(m.train_step(x, y))
        g1 = m.get_grads()
        losses = [l1] + [float(m.train_step(x, y)) for _ in range(3)]        # eager, capture + replay, replay
        out[rank] = dict(l1=l1, g1=g1, avg=avg, solo_avg=solo_avg, mean_loss=float(np.mean([p['loss'] for p in per])), losses=losses,
                         w=m.get_weights(), nlog=len(m._dp.launch_log), frontier=[f for f, _ in m._dp.launch_log])
        m.close()
    finally:
        dist.destroy_process_group()
        faulthandler.cancel_dump_traceback_later()


def test_two_rank_multiresunet_step_matches_oracle_subbatch_average():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (gpurun --gpus 2)')
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_multires, args=(world, _free_port(), out), nprocs=world, join=True)
    a, b = out[0], out[1]
    names = list(a['avg'])
    got = np.concatenate([a['g1'][k].ravel() for k in names])
    want = np.concatenate([a['avg'][k].ravel() for k in names])
    assert _rel(got, want) <= 2e-2, _rel(got, want)
    solo = np.concatenate([a['solo_avg'][k].ravel() for k in names])
    assert _rel(got, solo) <= 1e-5, _rel(got, solo)
    for k in names:
        np.testing.assert_array_equal(a['g1'][k], b['g1'][k])
    assert abs(a['l1'] - a['mean_loss']) <= 1e-4 * abs(a['mean_loss'])
    assert a['losses'] == b['losses']
    trainable = set(names)
    for k in a['w']:
        if k in trainable:
            np.testing.assert_array_equal(a['w'][k], b['w'][k])              # replicas stay mirrored (moving statistics are rank-local)
    assert a['nlog'] >= 4 and all(f == 0 for f in a['frontier'])             # every bucket left after the gather (dp.finish)


def _worker_engine(rank, world, port, save_path, out):
    """engine.TFKerasModel under ``deploy_options.enable_multigpu`` (multigpu.yaml): one process per GPU instead of the
    reference's MirroredStrategy; rank 0 writes the checkpoints, every rank resumes from them."""
    import faulthandler
    import warnings
    import torch.distributed as dist
    logdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    os.makedirs(logdir, exist_ok=True)
    log = open(os.path.join(logdir, f'dp_test_engine_rank{rank}.log'), 'w')
    faulthandler.dump_traceback_later(120, file=log, exit=True)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from dnncancerannotator_b200 import engine as E
        from dnncancerannotator_b200.synthetic import make_slices
        cfg = dict(model='UNetAnnotator', model_options=OPTS,
                   deploy_options=dict(optimizer='adam', loss=dict(class_name='WeightedCrossentropy', config=LOSS),
                                       LearningRateScheduler='lambda epoch, current_lr: 0.001', enable_multigpu=True))
        train = [make_slices(3, 32, 32, 3, seed=500 + 10 * i + rank) for i in range(2)]      # every rank its own shard
        val = [make_slices(2, 32, 32, 3, seed=900)]
        eng = E.TFKerasModel(cfg, dtype='fp32')
        h1 = eng.train(train, val_data=val, save_path=save_path, save_freq=2, max_steps=4)
        dist.barrier()
        ck = list(eng.get_ckpts(os.path.join(save_path, 'checkpoints')))
        eng2 = E.TFKerasModel(cfg, dtype='fp32')
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            h2 = eng2.train(train, val_data=val, save_path=save_path, save_freq=2, max_steps=6)
        out[rank] = dict(ck=ck, loss1=h1.history['loss'], val1=h1.history['val_loss'], epoch2=h2.epoch, loss2=h2.history['loss'],
                         step=eng2.current_step, w=eng2.model.get_weights(), world=eng.model._dp.world_size)
        eng.model.close()
        eng2.model.close()
    finally:
        dist.destroy_process_group()
        faulthandler.cancel_dump_traceback_later()


def test_two_rank_engine_train_checkpoint_resume(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (gpurun --gpus 2)')
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_engine, args=(2, _free_port(), str(tmp_path / 'run'), out), nprocs=2, join=True)
    a, b = out[0], out[1]
    assert a['world'] == 2 and a['ck'] == [2, 4] and b['ck'] == [2, 4]
    assert a['loss1'] == b['loss1']                        # the reported loss is the global mean on every rank
    assert a['step'] == 4 and b['step'] == 4 and a['epoch2'] == [4, 5] and a['loss2'] == b['loss2']
    files = sorted(os.listdir(tmp_path / 'run' / 'checkpoints'))
    assert files == sorted(f'ckpt-{s}.{e}' for s in (2, 4, 6) for e in ('index', 'data-00000-of-00001'))       # rank 0 alone wrote
    trainable = [k for k in a['w'] if not k.endswith(('moving_mean', 'moving_var'))]
    for k in trainable:
        np.testing.assert_array_equal(a['w'][k], b['w'][k])
