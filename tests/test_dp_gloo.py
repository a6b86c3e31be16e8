"""Data-parallel host logic on CPU: world_size-2 gloo run of the gradient all-reduce helper
(the same code path NCCL takes on the GPUs), bucket partitioning, parameter broadcast, and the
reference's DP semantics: loss scaled by 1/world + SUM all-reduce == averaging the
per-replica gradients of independent sub-batches (rank-local BN statistics and loss weight)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dnncancerannotator_b200.parallel import GradAllReduce, bucket_ranges


def test_bucket_ranges_cover_in_reverse_order():
    r = bucket_ranges(10, 4)
    assert r == [(6, 10), (2, 6), (0, 2)]
    assert bucket_ranges(8, 100) == [(0, 8)]
    covered = sorted(i for a, b in bucket_ranges(1001, 64) for i in range(a, b))
    assert covered == list(range(1001))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import ref_models as rm
        from dnncancerannotator_b200.synthetic import make_slices
        opts = dict(n_filters_first=4, n_downsample=2, rate=2, kernel_size=3, conv_stride=1, bn=True, padding='same')
        torch.set_num_threads(1)
        ref = rm.build_model('UNetAnnotator', opts, (None, 16, 16, 3), seed=rank)     # replicas start different ...
        dp = GradAllReduce(bucket_bytes=256)                                            # tiny buckets -> many buckets
        flat = torch.cat([ref.weights[k].reshape(-1) for k in ref.weights])
        dp.broadcast_parameters(flat)                                                   # ... and are mirrored from rank 0
        off = 0
        for k in ref.weights:
            n = ref.weights[k].numel()
            ref.weights[k] = flat[off:off + n].view_as(ref.weights[k]).clone()
            off += n
        x, y = make_slices(2, 16, 16, 3, seed=1234 + rank)                               # this rank's share of the batch
        r = ref.train_step_grads(x, y, dict(weight_mul=3.0), n_replicas=world)           # loss pre-scaled by 1/world
        g = torch.cat([r['grads'][k].reshape(-1) for k in ref.trainable]).contiguous()
        local = g.clone()
        dp.all_reduce(g)                                                                 # SUM over ranks, bucketed
        out[rank] = dict(local=local.numpy(), reduced=g.numpy(), w0=flat.numpy())
    finally:
        dist.destroy_process_group()


def test_allreduce_world2_gloo_matches_gradient_averaging():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    a, b = out[0], out[1]
    np.testing.assert_array_equal(a['w0'], b['w0'])                      # mirrored variables
    np.testing.assert_allclose(a['reduced'], b['reduced'], rtol=0, atol=0)
    np.testing.assert_allclose(a['reduced'], a['local'] + b['local'], rtol=1e-6, atol=1e-9)
    assert np.abs(a['local'] - b['local']).max() > 0                     # different sub-batches
