"""Pinned host staging buffers for the input tail (data.py:110 ``prefetch`` -> H2D).

One process per GPU copies its batch from page-locked host memory every step; with eight ranks on one
host the host memory system, not the GPUs, bounds the end-to-end rate (SCALE_r01: 0.69 efficiency at 8).
Two knobs of the allocation matter for a buffer the CPU only WRITES (the data loader) and the GPU only reads:

* ``cudaHostAllocWriteCombined``: the pages are mapped write-combining, so a DMA read of them is not snooped
  through the CPU caches (CUDA documents up to 40 % faster PCIe reads); CPU reads of such memory are slow, so it is
  for staging buffers only;
* NUMA placement: ``interleave=True`` sets ``MPOL_INTERLEAVE`` over all memory nodes for the duration of the
  allocation, so eight ranks do not all pull from the node the launcher happened to start on.

Buffers come from ``dnnca_host_alloc`` (libdnnca, ``cudaHostAlloc``) and are exposed as torch CPU tensors that
``Tensor.copy_(..., non_blocking=True)`` recognises as pinned.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import native as N

_LIVE = []          # (ptr, nbytes): freed at interpreter exit by the CUDA context teardown


def _numa_nodes():
    try:
        return sorted(int(d[4:]) for d in os.listdir('/sys/devices/system/node') if d.startswith('node') and d[4:].isdigit())
    except OSError:
        return [0]


def _set_mempolicy(mode, nodes):
    """set_mempolicy(2) through libc's syscall(); returns True when the kernel accepted it."""
    try:
        libc = C.CDLL(None, use_errno=True)
        mask = 0
        for n in nodes:
            mask |= 1 << n
        maxnode = max(nodes) + 2 if nodes else 0
        m = C.c_ulong(mask)
        SYS_set_mempolicy = 238          # x86_64
        return libc.syscall(SYS_set_mempolicy, C.c_int(mode), C.byref(m) if nodes else None, C.c_ulong(maxnode)) == 0
    except Exception:
        return False


def pinned_empty(shape, dtype=np.uint8, write_combined=True, interleave=None):
    """A page-locked host array (numpy view + torch tensor sharing it) from ``cudaHostAlloc``."""
    if interleave is None:
        interleave = os.environ.get('DNNCA_HOST_INTERLEAVE', '0') == '1'
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    nodes = _numa_nodes()
    policy = interleave and len(nodes) > 1 and _set_mempolicy(3, nodes)      # MPOL_INTERLEAVE
    try:
        p = C.c_void_p()
        N.check(N.lib().dnnca_host_alloc(C.c_size_t(max(nbytes, 16)), 1 if write_combined else 0, C.byref(p)), 'dnnca_host_alloc')
    finally:
        if policy:
            _set_mempolicy(0, [])                                            # MPOL_DEFAULT
    _LIVE.append((p.value, nbytes))
    buf = (C.c_uint8 * max(nbytes, 16)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    t = torch.from_numpy(arr)
    t._dnnca_keepalive = buf
    return t


def pinned_like(array, write_combined=True, interleave=None):
    """Copies ``array`` into a fresh pinned (optionally write-combined / NUMA-interleaved) buffer."""
    a = array.numpy() if torch.is_tensor(array) else np.ascontiguousarray(array)
    t = pinned_empty(a.shape, a.dtype, write_combined, interleave)
    t.numpy()[...] = a             # CPU writes only: fine for write-combined pages
    return t


def describe(t, write_combined=True):
    nodes = _numa_nodes()
    return dict(kind='cudaHostAlloc' + ('WriteCombined' if write_combined else 'Default'), numa_nodes=len(nodes),
                interleaved=os.environ.get('DNNCA_HOST_INTERLEAVE', '0') == '1' and len(nodes) > 1,
                is_pinned=bool(t.is_pinned()) if torch.cuda.is_available() else None)
