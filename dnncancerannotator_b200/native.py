"""ctypes binding of ``libdnnca.so`` (``include/dnnca.h``).

This is the only door between the Python host code and the CUDA kernels.  There
is **no CPU fallback**: if the shared library has not been built (run
``python -m dnncancerannotator_b200.build``) importing :func:`lib` raises, and
every call that returns a non-zero status raises ``DnncaError`` with the
library's message.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdnnca.so')

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2


class DnncaError(RuntimeError):
    pass


class Tensor(C.Structure):
    """``dnnca_tensor_t``: channel-slice view of an NHWC buffer."""
    _fields_ = [('data', C.c_void_p), ('n', C.c_int32), ('h', C.c_int32), ('w', C.c_int32), ('c', C.c_int32),
                ('cstride', C.c_int32), ('coff', C.c_int32), ('dtype', C.c_int32)]


class LabelStats(C.Structure):
    _fields_ = [('sum', C.c_double), ('min_key', C.c_uint32), ('max_key', C.c_uint32)]


class LossConfig(C.Structure):
    _fields_ = [('weight', C.c_float), ('has_weight', C.c_int32), ('weight_add', C.c_float),
                ('weight_mul', C.c_float), ('grad_scale', C.c_float)]


_TP = C.POINTER(Tensor)
_vp, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64

# name -> argtypes (all return int unless listed in _RESTYPES)
_PROTOS = {
    'dnnca_version': [],
    'dnnca_last_error': [],
    'dnnca_sm_count': [C.POINTER(C.c_int)],
    'dnnca_debug_force_generic': [_i],
    'dnnca_debug_launch_count': [_i],
    'dnnca_debug_family_count': [_i, _i],
    'dnnca_conv_workspace_bytes': [_i, _i, _i],
    'dnnca_conv2d_prepack': [_vp, _TP, _TP, _vp, _TP, _i, _vp, C.c_size_t],
    'dnnca_convtranspose2x2_prepack': [_vp, _TP, _vp, _TP, _vp, C.c_size_t],
    'dnnca_conv2d_fold_supported': [_TP, _TP, _TP, _i],
    'dnnca_conv2d_fold_scratch_bytes': [_i, _i],
    'dnnca_conv2d_fprop_affine': [_vp, _TP, _TP, _vp, _vp, _vp, _vp, _TP, _i, _f, _vp, _vp, C.c_size_t, _vp],
    'dnnca_conv2d_wgrad_affine': [_vp, _TP, _TP, _vp, _vp, _TP, _vp, _vp, _vp],
    'dnnca_maxpool2x2_fwd_affine': [_vp, _TP, _vp, _TP, _vp, _vp],
    'dnnca_conv2d_fprop': [_vp, _TP, _TP, _vp, _vp, _TP, _i, _i, _f, _vp, _vp, C.c_size_t],
    'dnnca_conv2d_dgrad': [_vp, _TP, _vp, _TP, _TP, _i, _TP, _i, _f, _vp, C.c_size_t],
    'dnnca_conv2d_dgrad_bnreduce': [_vp, _TP, _vp, _TP, _TP, _i, _TP, _vp, _vp, _vp, C.c_size_t],
    'dnnca_convtranspose2x2_dgrad_bnreduce': [_vp, _TP, _vp, _TP, _TP, _vp, _vp, _vp, C.c_size_t],
    'dnnca_conv2d_wgrad': [_vp, _TP, _TP, _TP, _vp, _vp, _i],
    'dnnca_convtranspose2x2_fprop': [_vp, _TP, _vp, _vp, _TP, _vp, _vp, C.c_size_t],
    'dnnca_convtranspose2x2_dgrad': [_vp, _TP, _vp, _TP, _TP, _i, _f, _vp, C.c_size_t],
    'dnnca_convtranspose2x2_wgrad': [_vp, _TP, _TP, _vp, _vp],
    'dnnca_maxpool2x2_fwd': [_vp, _TP, _TP, _vp, _vp],
    'dnnca_maxpool2x2_bwd': [_vp, _TP, _vp, _TP, _TP, _TP, _i, _f],
    'dnnca_channel_stats': [_vp, _TP, _vp],
    'dnnca_bn_finalize': [_vp, _vp, _i64, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp],
    'dnnca_bn_inference_params': [_vp, _i, _vp, _vp, _f, _vp, _vp, _vp],
    'dnnca_bn_apply': [_vp, _TP, _vp, _TP],
    'dnnca_fold_weights': [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _f, _vp, _vp],
    'dnnca_bn_inference_params_mapped': [_vp, _i, _vp, _vp, _vp, _f, _vp, _vp, _vp],
    'dnnca_bn_bwd_reduce': [_vp, _TP, _TP, _vp, _vp],
    'dnnca_bn_bwd_apply': [_vp, _TP, _TP, _vp, _vp, _vp, _TP, _i, _f, _vp, _vp],
    'dnnca_gaussian_filter2d': [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp],
    'dnnca_label_stats_init': [_vp, _vp],
    'dnnca_label_stats': [_vp, _vp, _i64, _vp],
    'dnnca_label_stats_decode': [C.POINTER(LabelStats), C.POINTER(C.c_double), C.POINTER(C.c_float),
                                 C.POINTER(C.c_float)],
    'dnnca_head_fwd': [_vp, _TP, _vp, _vp, _vp, _vp],
    'dnnca_head_input_grad': [_vp, _TP, _vp, _vp, _TP, _i, _f],
    'dnnca_head_bce_fwd_bwd': [_vp, _TP, _vp, _vp, _vp, _vp, C.POINTER(LossConfig), _vp, _vp, _vp, _TP, _i, _f,
                               _vp, _vp],
    'dnnca_threshold_hist': [_vp, _vp, _vp, _i64, _vp, _i, _vp],
    'dnnca_resize_bilinear': [_vp, _vp, _i, _i, _i, _vp, _i, _i],
    'dnnca_grey_open': [_vp, _vp, _i, _i, _i, _i, _vp],
    'dnnca_connected_components': [_vp, _vp, _i, _i, _i, _vp],
    'dnnca_region_workspace_bytes': [_i, _i, _i, _i, _i],
    'dnnca_region_confusion': [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _f, _i, _vp, C.c_size_t, _i, _vp, _vp, _vp],
    'dnnca_add_relu_affine': [_vp, _TP, _vp, _TP, _vp, _vp, _TP],
    'dnnca_bn_apply_act': [_vp, _TP, _vp, _TP, _i, _f],
    'dnnca_act_bwd': [_vp, _TP, _TP, _TP, _i, _f],
    'dnnca_accumulate': [_vp, _TP, _TP],
    'dnnca_gather_f32': [_vp, _vp, _vp, _i64, _vp],
    'dnnca_head_conv_bwd': [_vp, _TP, _vp, _vp, _TP, _i, _f, _vp],
    'dnnca_u8_to_unit': [_vp, _vp, _i64, _vp, _i],
    'dnnca_input_tail': [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, C.POINTER(C.c_int32), _i, _i, _vp, _i, _i, _vp],
    'dnnca_unpack_label_bits': [_vp, _vp, _i64, _vp],
    'dnnca_tps_workspace_bytes': [_i, _i],
    'dnnca_tps_fit': [_vp, _vp, _vp, _i, _i, _f, _vp, C.c_size_t, _vp, _vp],
    'dnnca_tps_warp': [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _f, _vp, _vp],
    'dnnca_convert': [_vp, _TP, _TP],
    'dnnca_adam_step': [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp],
    'dnnca_p2p_alloc': [C.c_size_t, C.POINTER(C.c_void_p)],
    'dnnca_p2p_free': [_vp],
    'dnnca_p2p_export': [_vp, C.c_char_p],
    'dnnca_p2p_import': [C.c_char_p, C.POINTER(C.c_void_p)],
    'dnnca_p2p_close': [_vp],
    'dnnca_p2p_wait_done': [_vp, _vp, _i, _vp],
    'dnnca_p2p_adam_step': [_vp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _i, _i, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp],
    'dnnca_nccl_unique_id': [C.c_char_p],
    'dnnca_nccl_comm_init_rank': [C.POINTER(C.c_void_p), _i, C.c_char_p, _i],
    'dnnca_nccl_comm_destroy': [_vp],
    'dnnca_nccl_allreduce_bucket': [_vp, _vp, _vp, _i64, _i],
    'dnnca_nccl_broadcast': [_vp, _vp, _vp, _i64, _i],
    'dnnca_host_alloc': [C.c_size_t, _i, C.POINTER(C.c_void_p)],
    'dnnca_host_free': [_vp],
    'dnnca_loss_total': [_vp, _vp, _i, _vp, _vp, _i64, _f, _vp],
}
_RESTYPES = {'dnnca_last_error': C.c_char_p, 'dnnca_label_stats_decode': None,
             'dnnca_debug_launch_count': C.c_longlong, 'dnnca_debug_family_count': C.c_longlong,
             'dnnca_conv_workspace_bytes': C.c_size_t, 'dnnca_conv2d_fold_scratch_bytes': C.c_size_t,
             'dnnca_region_workspace_bytes': C.c_size_t, 'dnnca_tps_workspace_bytes': C.c_size_t}

_lib = None


def exported_symbols():
    """Every symbol ``include/dnnca.h`` declares (used by the CPU-side ABI test)."""
    return sorted(_PROTOS)


def lib():
    """Loads libdnnca.so once; raises if it is missing (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DnncaError(
                f'{LIB_PATH} is missing: build it with `python -m dnncancerannotator_b200.build` '
                '(nvcc, sm_100a). dnncancerannotator_b200 has no CPU or PyTorch fallback.')
        handle = C.CDLL(LIB_PATH)
        for name, argtypes in _PROTOS.items():
            fn = getattr(handle, name)          # AttributeError = ABI mismatch, fail loudly
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = handle
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = lib().dnnca_last_error()
        raise DnncaError(f'{what or "dnnca call"} failed ({rc}): {msg.decode() if msg else "?"}')


_profiler = None   # set by Profiler: records (name, args, start event, end event) per C-ABI call


def call(name, *args):
    if _profiler is not None:
        _profiler.before(name, args)
        check(getattr(lib(), name)(*args), name)
        _profiler.after()
        return
    check(getattr(lib(), name)(*args), name)


class Profiler:
    """Times every C-ABI call of an eager (non-graph) pass with CUDA events on the launching
    stream and estimates its algorithmic bytes / conv FLOPs from the tensor arguments
    (each activation read once per consumer and written once at its storage dtype)."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        global _profiler
        _profiler = self
        return self

    def __exit__(self, *a):
        global _profiler
        _profiler = None

    def before(self, name, args):
        self._cur = (name, self._cost(name, args), torch.cuda.Event(enable_timing=True),
                     torch.cuda.Event(enable_timing=True))
        self._cur[2].record(torch.cuda.current_stream())

    def after(self):
        self._cur[3].record(torch.cuda.current_stream())
        self.records.append(self._cur)

    @staticmethod
    def _views(args):
        out = []
        for a in args:
            obj = getattr(a, '_obj', None)
            if isinstance(obj, Tensor):
                out.append(obj)
        return out

    def _cost(self, name, args):
        vs = self._views(args)
        esz = lambda v: 4 if v.dtype == F32 else 2
        byt = sum(v.n * v.h * v.w * v.c * esz(v) for v in vs)
        flops = 0
        label = name.replace('dnnca_', '')
        if name in ('dnnca_conv2d_prepack', 'dnnca_conv2d_fold_supported'):
            return dict(label=label, bytes=0, flops=0)
        if name.endswith('_bnreduce'):              # dgrad + the BatchNorm backward sums in its epilogue: same conv
            name = name[:-len('_bnreduce')]
            label = name.replace('dnnca_', '')
        if name.startswith('dnnca_conv2d_') and len(vs) >= 2:
            affine = name.endswith('_affine')      # BatchNorm folded into the input(s): same conv, k = 3
            k = 3 if affine else [a for a in args if isinstance(a, int)][0]
            a = vs[0]
            name = name[:-len('_affine')] if affine else name
            label = name.replace('dnnca_', '')
            if name.endswith('fprop'):      # x [x2] y
                cin, cout = sum(v.c for v in vs[:-1]), vs[-1].c
            elif name.endswith('wgrad'):    # x [x2] dz
                cin, cout = sum(v.c for v in vs[:-1]), vs[-1].c
            else:                            # dz dx [dx2] [mask]: kernel-input = dz, kernel-output = dx (+dx2)
                nout = 2 if (len(args) > 4 and getattr(args[4], '_obj', None) is not None) else 1
                cin, cout = vs[0].c, sum(v.c for v in vs[1:1 + nout])
            flops = 2 * a.n * a.h * a.w * cin * cout * k * k
            byt += 4 * k * k * cin * cout
            label += f'[{cin}->{cout}@{a.h}]'
        elif name.startswith('dnnca_convtranspose2x2_') and len(vs) >= 2:
            a, b = vs[0], vs[1]
            small = a if a.h < b.h else b
            flops = 2 * small.n * small.h * small.w * a.c * b.c * 4
            byt += 16 * a.c * b.c
            label += f'[{a.c}->{b.c}@{small.h}]'
        elif name == 'dnnca_maxpool2x2_fwd' and vs:
            byt += vs[1].n * vs[1].h * vs[1].w * vs[1].c
            label += f'[{vs[0].c}@{vs[0].h}]'
        elif name == 'dnnca_maxpool2x2_bwd' and vs:
            byt += vs[0].n * vs[0].h * vs[0].w * vs[0].c
            label += f'[{vs[0].c}@{vs[0].h}]'
        elif name in ('dnnca_head_bce_fwd_bwd', 'dnnca_head_fwd') and vs:
            f = vs[0]
            byt += f.n * f.h * f.w * 4 * (3 if name.endswith('bwd') else 2)
            label += f'[{f.c}@{f.h}]'
        elif name == 'dnnca_label_stats':
            byt = 4 * [a for a in args if isinstance(a, int)][0]
        elif name == 'dnnca_adam_step':
            byt = 28 * [a for a in args if isinstance(a, int)][0]
        elif vs:
            label += f'[{vs[0].c}@{vs[0].h}]'
        return dict(label=label, bytes=byt, flops=flops)

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, cost, e0, e1 in self.records:
            d = agg.setdefault(cost['label'], dict(ms=0.0, calls=0, bytes=0, flops=0))
            d['ms'] += e0.elapsed_time(e1)
            d['calls'] += 1
            d['bytes'] += cost['bytes']
            d['flops'] += cost['flops']
        return agg


def stream_ptr():
    """cudaStream_t of torch's current stream (all launches go there)."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(dt):
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise DnncaError(f'unsupported activation dtype {dt}')


def ptr(t):
    """Device pointer of a torch tensor (or None) as c_void_p."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def tensor_view(buf: torch.Tensor, coff=0, c=None) -> Tensor:
    """``dnnca_tensor_t`` for channels [coff, coff+c) of a contiguous NHWC torch buffer."""
    assert buf.dim() == 4 and buf.is_contiguous(), 'NHWC contiguous buffer expected'
    n, h, w, cs = buf.shape
    c = cs - coff if c is None else c
    assert 0 <= coff and coff + c <= cs
    return Tensor(buf.data_ptr(), n, h, w, c, cs, coff, dtype_code(buf.dtype))
