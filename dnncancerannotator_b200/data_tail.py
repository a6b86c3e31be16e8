"""Device side of the reference's input pipeline tail (``annotator/data.py``).

``prepare_batch`` takes the RAW combined uint8 slices ``[B, Hin, Win, S]`` (every slice type incl. the label as one
channel -- what ``tf_prepare_combined_slices`` / the TFRecord reader yield before ``base`` touches them) and runs, in one
kernel (``dnnca_input_tail``): the centre crop of ``base`` (data.py:182-197), the displaced crop of ``random_crop``
(data.py:677-689), ``tf.image.random_flip_left_right`` (data.py:620-625), the float cast and ``/255`` (data.py:198-199)
and the feature / label split of ``to_feature_label`` (data.py:766-788).  The random draws stay on the host
(``draw_augmentation`` restates the reference's distributions); the pixels never exist as float32 on the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import native as N

DEFAULT_SLICE_TYPES = ('TRA', 'ADC', 'DWI', 'DCEE', 'DCEL', 'label')     # data.py:30-41 / data_options.yaml:7


def draw_augmentation(rng: np.random.Generator, batch, in_size, out_size, random_crop=True, random_flip=True,
                      stddev=4, max_=6, min_=-6):
    """Per-sample crop origins ``[B,2]`` (row, column) and flip flags ``[B]``.

    ``random_crop`` (data.py:677-689): origin = (in - out)//2 + clip(int(normal(0, stddev)), min_, max_) per axis
    (``tf.cast(..., tf.int32)`` truncates towards zero); ``random_flip_left_right`` flips with probability 1/2."""
    hin, win = in_size
    hout, wout = out_size
    base = np.array([(hin - hout) // 2, (win - wout) // 2], np.int64)
    if random_crop:
        diff = np.clip(np.trunc(rng.normal(0.0, stddev, (batch, 2))).astype(np.int64), min_, max_)
    else:
        diff = np.zeros((batch, 2), np.int64)
    origin = base[None, :] + diff
    origin[:, 0] = np.clip(origin[:, 0], 0, hin - hout)      # crop_to_bounding_box would raise outside the image
    origin[:, 1] = np.clip(origin[:, 1], 0, win - wout)
    flip = (rng.random(batch) < 0.5) if random_flip else np.zeros(batch, bool)
    return origin.astype(np.int32), flip.astype(np.uint8)


def prepare_batch(combined, slice_types=DEFAULT_SLICE_TYPES, output_size=(256, 256), crop_yx=None, flip=None,
                  dtype=torch.float32, x_out=None, y_out=None, device=None):
    """-> ``(x [B,H,W,nf] dtype, y [B,H,W] float32)`` on the device.

    ``combined``: uint8 ``[B,Hin,Win,S]`` (numpy, CPU or CUDA tensor; a pinned CPU tensor is copied asynchronously).
    ``crop_yx`` / ``flip``: outputs of ``draw_augmentation`` (None = centre crop / no flip = the evaluation pipeline,
    data.py:114-144).  ``x_out`` may be a wider pre-allocated buffer ``[B,H,W,Cpad]`` (e.g. the padded bf16 input of the
    first conv): only its first nf channels are written."""
    device = device or torch.device('cuda', torch.cuda.current_device())
    t = combined if torch.is_tensor(combined) else torch.from_numpy(np.ascontiguousarray(combined))
    if t.dtype != torch.uint8 or t.dim() != 4:
        raise TypeError('combined slices must be uint8 [B,Hin,Win,S]')
    t = t.to(device, non_blocking=True).contiguous()
    B, hin, win, S = t.shape
    slice_types = list(slice_types)
    if len(slice_types) != S:
        raise ValueError(f'{S} channels but {len(slice_types)} slice types')
    fidx = [i for i, n in enumerate(slice_types) if n != 'label']                     # data.py:770
    lidx = slice_types.index('label') if 'label' in slice_types else -1                # data.py:771
    hout, wout = output_size
    if x_out is None:
        x_out = torch.empty(B, hout, wout, len(fidx), dtype=dtype, device=device)
    if y_out is None:
        y_out = torch.empty(B, hout, wout, dtype=torch.float32, device=device)
    assert x_out.is_contiguous() and tuple(x_out.shape[:3]) == (B, hout, wout) and x_out.shape[3] >= len(fidx)
    cy = None if crop_yx is None else torch.as_tensor(np.ascontiguousarray(crop_yx, np.int32)).to(device)
    fl = None if flip is None else torch.as_tensor(np.ascontiguousarray(flip, np.uint8)).to(device)
    if cy is not None:
        c = np.asarray(crop_yx)
        if c.shape != (B, 2) or (c < 0).any() or (c[:, 0] > hin - hout).any() or (c[:, 1] > win - wout).any():
            raise ValueError('crop origins must be [B,2] and keep the window inside the image')
    arr = (C.c_int32 * len(fidx))(*fidx)
    N.call('dnnca_input_tail', N.stream_ptr(), N.ptr(t), B, hin, win, S, N.ptr(cy), N.ptr(fl), hout, wout, arr, len(fidx),
           lidx, N.ptr(x_out), N.dtype_code(x_out.dtype), x_out.shape[3], N.ptr(y_out))
    return x_out, y_out


# ---- binary labels as bits --------------------------------------------------------------------------------------------

class PackedLabels:
    """Binary label masks ``[B,H,W]`` as bits (``numpy.packbits`` along W): what ``pack_labels`` returns and
    ``Model.train_step`` / ``prefetch`` accept in place of ``y``.  The label channel of the reference's
    input contract is a PNG mask divided by 255 (data.py:193-206): 0 or 1, so one bit per pixel loses nothing and the
    host->device copy of a batch shrinks from 4 to 3.125 bytes per pixel (uint8 images + labels)."""

    def __init__(self, bits, shape):
        self.bits, self.shape = bits, tuple(shape)

    @property
    def nbytes(self):
        return int(self.bits.numel())


def pack_labels(y, pinned=True):
    """``y``: labels ``[B,H,W]``, float in {0, 1} or uint8 in {0, 255} (anything else raises: fractional labels, e.g. after
    the warp augmentation on the host, cannot be packed).  W must be a multiple of 8."""
    a = y.cpu().numpy() if torch.is_tensor(y) else np.asarray(y)
    if a.ndim != 3 or a.shape[2] % 8:
        raise ValueError('labels must be [B,H,W] with W a multiple of 8')
    one = 255 if a.dtype == np.uint8 else 1
    if not np.all((a == 0) | (a == one)):
        raise ValueError('only binary label masks can be packed')
    bits = torch.from_numpy(np.packbits(a != 0, axis=-1))
    if pinned and torch.cuda.is_available():
        from . import hostmem
        bits = hostmem.pinned_like(bits, write_combined=False)
    return PackedLabels(bits, a.shape)


# ---- thin-plate-spline warp augmentation (data.py:628-645, 718-763) ---------------------------------------------------

def draw_control_points(rng: np.random.Generator, n_images, width, n_points=100, max_diff=5, stddev=2.0):
    """``random_warp``'s random draws (data.py:742-746): source points ~ U[0, width)^2, displacement ~ N(0, stddev)
    clipped to +-max_diff.  -> (source, dest) float32 ``[n_images, n_points, 2]`` in (row, column) pixels."""
    raw = rng.uniform(0.0, float(width), (n_images, n_points, 2)).astype(np.float32)
    diff = np.clip(rng.normal(0.0, stddev, (n_images, n_points, 2)), -float(max_diff), float(max_diff)).astype(np.float32)
    return raw, raw + diff


def sparse_image_warp(image, source_control_point_locations, dest_control_point_locations, interpolation_order=2,
                      regularization_weight=0.0, num_boundary_points=0, return_flow=True):
    """``tfa.image.sparse_image_warp`` as the reference calls it (data.py:749-753) -> ``(warped image, dense flow)``.

    ``image``: float32 ``[B,H,W,C]`` (numpy / CPU / CUDA tensor); control points ``[B,P,2]`` (row, column).  Only the
    arguments the reference uses are implemented (thin-plate order 2, no regularisation, no boundary points); the fit
    and the dense flow run in FP64 in normalised coordinates (``csrc/tps_warp.cu``)."""
    if interpolation_order != 2 or regularization_weight != 0.0 or num_boundary_points != 0:
        raise NotImplementedError('sparse_image_warp: the reference uses interpolation_order=2, regularization_weight=0, '
                                  'num_boundary_points=0 (data.py:749-753); nothing else is built')
    dev = image.device if torch.is_tensor(image) and image.is_cuda else torch.device('cuda', torch.cuda.current_device())
    img = torch.as_tensor(image).to(device=dev, dtype=torch.float32).contiguous()
    src = torch.as_tensor(source_control_point_locations).to(device=dev, dtype=torch.float32).contiguous()
    dst = torch.as_tensor(dest_control_point_locations).to(device=dev, dtype=torch.float32).contiguous()
    if img.dim() != 4 or src.shape != dst.shape or src.dim() != 3 or src.shape[0] != img.shape[0] or src.shape[2] != 2:
        raise ValueError('image [B,H,W,C]; control points [B,P,2]')
    B, H, W, Cc = img.shape
    P = src.shape[1]
    lib = N.lib()
    ws = torch.empty(int(lib.dnnca_tps_workspace_bytes(B, P)) // 8 + 1, dtype=torch.float64, device=dev)
    coef = torch.empty(B, P + 3, 2, dtype=torch.float64, device=dev)
    singular = torch.zeros(1, dtype=torch.int32, device=dev)
    extent = float(max(H, W))
    N.call('dnnca_tps_fit', N.stream_ptr(), N.ptr(src), N.ptr(dst), B, P, extent, N.ptr(ws), ws.numel() * 8, N.ptr(coef),
           N.ptr(singular))
    out = torch.empty_like(img)
    flow = torch.empty(B, H, W, 2, dtype=torch.float32, device=dev) if return_flow else None
    N.call('dnnca_tps_warp', N.stream_ptr(), N.ptr(img), B, H, W, Cc, N.ptr(dst), N.ptr(coef), P, extent, N.ptr(out), N.ptr(flow))
    if int(singular.item()):
        raise N.DnncaError('sparse_image_warp: singular spline system (coincident control points)')
    return out, flow


def random_warp(image, n_points=100, max_diff=5, stddev=2.0, process_in_batch=None, rng=None):
    """``random_warp`` (data.py:718-763): a single image ``[H,W,C]`` (``process_in_batch=None``) or a batch
    ``[process_in_batch,H,W,C]``, square images only (data.py:741).  Random draws come from ``rng`` on the host."""
    rng = rng or np.random.default_rng()
    img = torch.as_tensor(image)
    single = process_in_batch is None
    if single:
        img = img[None]
    elif img.shape[0] != process_in_batch:
        raise ValueError(f'batch of {img.shape[0]} images but process_in_batch={process_in_batch}')
    if img.shape[1] != img.shape[2]:
        raise ValueError('random_warp: only square images are supported (data.py:741)')
    src, dst = draw_control_points(rng, img.shape[0], img.shape[1], n_points, max_diff, stddev)
    out, _ = sparse_image_warp(img, src, dst, return_flow=False)
    return out[0] if single else out
