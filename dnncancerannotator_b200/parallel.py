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

    def launch_ready(self, frontier, before=None):
        """Issues every not-yet-issued bucket that lies entirely in [frontier, end).  ``before()`` runs once ahead of
        the first bucket issued by this call (the step joins its side streams there: a bucket must not leave before the
        weight-gradient kernels that fill it)."""
        if self.world_size == 1:
            return
        if before is not None and self._next < len(self._buckets) and self._buckets[self._next][0] >= frontier:
            before()
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


class _DevMem:
    """Zero-copy torch view of raw device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr, nbytes, typestr, shape):
        self.__cuda_array_interface__ = dict(shape=shape, typestr=typestr, data=(int(ptr), False), version=3, strides=None)
        self.ptr, self.nbytes = ptr, nbytes


class P2PAdam:
    """All-reduce fused into Adam over NVLink peer memory (``csrc/p2p_adam.cu``) for models whose whole gradient is small
    (configs/unet.yaml: 35 KB; configs/mulmo_unet.yaml: 6.9 MB): the flat gradient buffer of every rank is peer-mapped through CUDA IPC and ONE kernel per
    rank sums the peers' gradients in place of an NCCL all-reduce and applies the Adam update.  ``torch.distributed`` is
    used only to exchange the 64-byte IPC handles at set-up."""

    MAX_BYTES = 8 << 20      # mulmo_unet.yaml (6.9 MB) included: measured faster than the overlapped NCCL buckets

    def __init__(self, group=None):
        import ctypes as C
        from . import native as N
        self.N, self.C = N, C
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.local = self.peers_g = self.peers_f = None
        self.opened = []

    @classmethod
    def usable(cls, nbytes, world):
        import os
        limit = int(float(os.environ.get('DNNCA_P2P_MAX_MB', cls.MAX_BYTES / (1 << 20))) * (1 << 20))
        return os.environ.get('DNNCA_P2P', '1') != '0' and 1 < world <= 8 and nbytes <= limit and torch.cuda.is_available()

    def setup(self, numel, device):
        """Allocates this rank's shared gradient buffer [numel] fp32 + flag array, exchanges IPC handles, opens the peers'.
        Returns the gradient buffer as a torch tensor (to replace ParamStore.grads_full)."""
        N, C = self.N, self.C
        lib = N.lib()
        gp, fp = C.c_void_p(), C.c_void_p()
        N.check(lib.dnnca_p2p_alloc(C.c_size_t(numel * 4), C.byref(gp)), 'p2p_alloc')
        N.check(lib.dnnca_p2p_alloc(C.c_size_t(256), C.byref(fp)), 'p2p_alloc')
        hg, hf = C.create_string_buffer(64), C.create_string_buffer(64)
        N.check(lib.dnnca_p2p_export(gp, hg), 'p2p_export')
        N.check(lib.dnnca_p2p_export(fp, hf), 'p2p_export')
        handles = [None] * self.world_size
        dist.all_gather_object(handles, (hg.raw, hf.raw), group=self.group)
        G, Fl = (C.c_void_p * self.world_size)(), (C.c_void_p * self.world_size)()
        for r, (a, b) in enumerate(handles):
            if r == self.rank:
                G[r], Fl[r] = gp.value, fp.value
                continue
            pa, pb = C.c_void_p(), C.c_void_p()
            N.check(lib.dnnca_p2p_import(a, C.byref(pa)), 'p2p_import')
            N.check(lib.dnnca_p2p_import(b, C.byref(pb)), 'p2p_import')
            self.opened += [pa, pb]
            G[r], Fl[r] = pa.value, pb.value
        self.peers_g, self.peers_f = G, Fl
        self.local = (gp, fp)
        self.numel = numel
        self.grads = torch.as_tensor(_DevMem(gp.value, numel * 4, '<f4', (numel,)), device=device)
        self.flags = torch.as_tensor(_DevMem(fp.value, 256, '<i8', (32,)), device=device)
        self.reduced = torch.zeros(numel, dtype=torch.float32, device=device)
        self.epoch = torch.zeros(1, dtype=torch.int64, device=device)
        self.done_blocks = torch.zeros(1, dtype=torch.int32, device=device)
        dist.barrier(group=self.group)        # every peer mapping exists before the first exchange
        return self.grads

    def wait_done(self):
        N = self.N
        N.call('dnnca_p2p_wait_done', N.stream_ptr(), self.local[1], self.world_size, N.ptr(self.epoch))

    def adam_step(self, ps):
        N = self.N
        N.call('dnnca_p2p_adam_step', N.stream_ptr(), self.peers_g, self.peers_f, self.rank, self.world_size, N.ptr(ps.params),
               N.ptr(ps.m), N.ptr(ps.v), ps.n_trainable, self.numel, N.ptr(self.reduced), N.ptr(ps.hyper), N.ptr(ps.step),
               N.ptr(self.epoch), N.ptr(ps.l2), N.ptr(self.done_blocks))

    def check(self):
        """Raises if a peer wait timed out (sticky device flag)."""
        if int(self.flags[16]) != 0:
            raise RuntimeError('p2p_adam: a wait for a peer GPU timed out (a rank died or fell out of lockstep)')

    def close(self):
        lib = self.N.lib()
        torch.cuda.synchronize()
        for p in self.opened:
            lib.dnnca_p2p_close(p)
        self.opened = []
