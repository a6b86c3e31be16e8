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
    """SUM all-reduce of the flat fp32 gradient buffer, bucketed in reverse layer order and OVERLAPPED with the
    backward pass: ``begin`` / ``launch_ready`` / ``finish`` are called from inside the training-step launch
    sequence (and therefore from inside its CUDA-graph capture -- NCCL collectives are capturable), each bucket is
    issued with ``async_op=True`` the moment the backward pass has written its last gradient, so NCCL's stream forks
    off the compute stream there and joins it again in ``finish`` right before the fused Adam."""

    def __init__(self, process_group=None, bucket_bytes=8 << 20, min_buckets=4):
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError('torch.distributed is not initialised (launch with torchrun / init_process_group)')
        self.group = process_group
        self.world_size = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        self.bucket_elems = max(bucket_bytes // 4, 1)
        self.min_buckets = max(int(min_buckets), 1)
        self._flat = None
        self._buckets, self._works, self._next = [], [], 0
        self.launch_log = []           # (frontier, bucket) per launch of the last step: tests / DESIGN evidence

    def buckets_for(self, numel):
        """Reverse-order buckets of at most ``bucket_bytes``; small models are still cut into ``min_buckets`` pieces
        so that all but the last piece travel while the backward pass is still running."""
        elems = min(self.bucket_elems, max((numel + self.min_buckets - 1) // self.min_buckets, 1))
        elems = (elems + 3) // 4 * 4
        return bucket_ranges(numel, elems)

    # ---- overlapped protocol ----------------------------------------------------------------------
    def begin(self, flat_grads: torch.Tensor):
        self._flat = flat_grads
        self._buckets = self.buckets_for(flat_grads.numel())
        self._works, self._next = [], 0
        self.launch_log = []

    def launch_ready(self, frontier):
        """Issues every not-yet-issued bucket that lies entirely in [frontier, end)."""
        if self.world_size == 1:
            return
        while self._next < len(self._buckets) and self._buckets[self._next][0] >= frontier:
            a, b = self._buckets[self._next]
            self._works.append(dist.all_reduce(self._flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self.launch_log.append((int(frontier), (a, b)))
            self._next += 1

    def finish(self):
        self.launch_ready(0)
        for w in self._works:
            w.wait()
        self._works = []

    # ---- one-shot form ----------------------------------------------------------------------------
    def all_reduce(self, flat_grads: torch.Tensor):
        """In-place SUM over ranks of a flat fp32 gradient buffer (no overlap)."""
        if self.world_size == 1:
            return flat_grads
        self.begin(flat_grads)
        self.finish()
        return flat_grads

    def broadcast_parameters(self, *flat_buffers, src=0):
        """Mirrored variables start identical on every replica."""
        for t in flat_buffers:
            dist.broadcast(t, src=src, group=self.group)

    def average(self, *flat_buffers):
        """Mean over replicas (BatchNorm moving statistics are aggregated this way when read / saved under
        MirroredStrategy [TF-semantics])."""
        for t in flat_buffers:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(self.world_size)
