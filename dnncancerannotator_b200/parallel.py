"""Synchronous data parallelism for the training step (``engine.py:260-263``:
``tf.distribute.MirroredStrategy`` -> one SUM all-reduce of all gradients per step).

One process per GPU (torchrun / ``torch.distributed``): every rank holds the full
parameters, takes its share of the global batch, keeps BatchNorm statistics and the
loss's positive-rate weight rank-local (reference behaviour), scales its
loss by 1/world (``dnnca_loss_config_t.grad_scale``) and SUM-all-reduces the flat fp32
gradient buffer.  The reduction is issued in buckets so that NCCL (NVLink 5 / NVSwitch)
pipelines them; the ``gloo`` backend runs the same code on CPU tensors for tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def bind_host_to_gpu(device_index=None):
    """Pin this process to the CPUs NVML reports as local to its GPU, so the pinned staging buffers it allocates afterwards
    (first touch) and the threads that fill them sit on the GPU's NUMA node: with one rank per GPU every rank otherwise
    allocates from the socket the launcher happened to start on and all host->device copies cross one memory controller.
    Returns the CPU list it bound to, or None when NVML / the affinity call is unavailable (nothing changes then)."""
    import os
    try:
        import pynvml
        idx = torch.cuda.current_device() if device_index is None else device_index
        props = torch.cuda.get_device_properties(idx)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + str(props.uuid)).encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(local & allowed)
        if not cpus or len(cpus) == len(allowed):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:       # no NVML, no uuid, restricted cpuset: keep the inherited affinity
        return None


def bucket_ranges(numel, bucket_elems):
    """[(start, stop)] covering [0, numel) in reverse order (last layers' gradients are ready
    first during the backward pass)."""
    bucket_elems = max(int(bucket_elems), 1)
    stops = list(range(numel, 0, -bucket_elems))
    return [(max(s - bucket_elems, 0), s) for s in stops]


class GradAllReduce:
    def __init__(self, process_group=None, bucket_bytes=8 << 20):
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError('torch.distributed is not initialised (launch with torchrun / init_process_group)')
        self.group = process_group
        self.world_size = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        self.bucket_elems = max(bucket_bytes // 4, 1)

    def all_reduce(self, flat_grads: torch.Tensor):
        """In-place SUM over ranks of a flat fp32 gradient buffer."""
        if self.world_size == 1:
            return flat_grads
        works = []
        for a, b in bucket_ranges(flat_grads.numel(), self.bucket_elems):
            works.append(dist.all_reduce(flat_grads[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in works:
            w.wait()
        return flat_grads

    def broadcast_parameters(self, *flat_buffers, src=0):
        """Mirrored variables start identical on every replica."""
        for t in flat_buffers:
            dist.broadcast(t, src=src, group=self.group)
