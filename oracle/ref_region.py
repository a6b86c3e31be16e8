"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the reference's REGION-BASED metrics
(annotator/utils/metrics.py:80-520, annotator/utils/image.py:12-29), SURVEY 8f "later" row.

PARITY PINNED for this module: the reference ships known-answer tests for exactly this path
(annotator/tests/test_region_metrics.py); ``tests/test_region_metrics.py`` restates every one of them (same sample
generators, same expected counts by construction) and runs them against this oracle (CPU) and against the CUDA path
(``-m gpu``).  The third-party pieces the reference calls are restated from their published behaviour and are
cross-checked against independent implementations available here:

* ``tfa.image.connected_components`` (tensorflow-addons, unpinned in requirements.txt:3): 4-connectivity, ids 1..n
  consecutive across the whole batch, ordered by the first pixel of a component in (image, row, column) order; zero
  stays 0.  Cross-check: ``scipy.ndimage.label`` (default cross structure) per image.
* ``tf.nn.erosion2d`` / ``tf.nn.dilation2d`` with an all-zero structuring element and 'SAME' padding
  (image.py:20-28): min / max over the k x k window, positions outside the image do not take part.
  Cross-check: ``scipy.ndimage.minimum_filter`` / ``maximum_filter`` with a neutral constant border.
* ``tf.image.resize`` (TF 2 default: bilinear, half-pixel centres, no antialias) (metrics.py:196-204).
  Cross-check: ``torch.nn.functional.interpolate(mode='bilinear', align_corners=False)``.

Everything here is written for clarity, not speed: it materialises the one-hot region masks and the
[labels, predictions, H, W, thresholds] broadcast exactly as the reference does.
"""
import numpy as np


# ---- third-party ops restated ------------------------------------------------------------------------------------

def connected_components(images):
    """tfa.image.connected_components for a batch [N,H,W] (bool / int): int32 ids, see module docstring."""
    images = np.asarray(images)
    assert images.ndim == 3
    n, h, w = images.shape
    out = np.zeros((n, h, w), np.int32)
    next_id = 0
    for b in range(n):
        img = images[b]
        for y in range(h):
            for x in range(w):
                if img[y, x] == 0 or out[b, y, x] != 0:
                    continue
                next_id += 1
                v = img[y, x]
                out[b, y, x] = next_id
                stack = [(y, x)]
                while stack:
                    cy, cx = stack.pop()
                    for ny, nx in ((cy - 1, cx), (cy + 1, cx), (cy, cx - 1), (cy, cx + 1)):
                        if 0 <= ny < h and 0 <= nx < w and out[b, ny, nx] == 0 and img[ny, nx] == v:
                            out[b, ny, nx] = next_id
                            stack.append((ny, nx))
    return out


def _window_reduce(x, k, fn):
    """fn (np.min / np.max) over the k x k 'SAME' window, out-of-image positions skipped ([TF-semantics]:
    dilation2d starts from the lowest value and only visits in-bounds taps; erosion2d = -dilation2d(-x, reversed kernel))."""
    n, h, w = x.shape
    before = (k - 1) // 2
    out = np.empty_like(x)
    for y in range(h):
        y0, y1 = max(y - before, 0), min(y - before + k, h)
        for xx in range(w):
            x0, x1 = max(xx - before, 0), min(xx - before + k, w)
            out[:, y, xx] = fn(x[:, y0:y1, x0:x1], axis=(1, 2))
    return out


def morph_open(image, filter_size):
    """image.py:12-29 for [N,H,W,1] (or [N,H,W]) integer / float images: erosion then dilation, zero structuring element."""
    x = np.asarray(image)
    squeeze = x.ndim == 4
    if squeeze:
        assert x.shape[-1] == 1
        x = x[..., 0]
    opened = _window_reduce(_window_reduce(x, filter_size, np.min), filter_size, np.max)
    return opened[..., None] if squeeze else opened


def resize_target(size, resize_factor):
    """metrics.py:199-200: tf.cast(tf.cast(size, tf.float16) * resize_factor, tf.int32) (the product is a float16)."""
    return int(np.float16(np.float16(size) * np.float16(resize_factor)))


def resize_bilinear(image, out_h, out_w):
    """tf.image.resize(image, [out_h, out_w]) for [N,H,W,C] float32 ([TF-semantics]: bilinear, half_pixel_centers,
    antialias=False; weights computed in float32 like the kernel's compute_interpolation_weights)."""
    img = np.asarray(image, np.float32)
    n, h, w, c = img.shape

    def weights(out_size, in_size):
        scale = np.float32(in_size) / np.float32(out_size)
        i = np.arange(out_size, dtype=np.float32)
        src = (i + np.float32(0.5)) * scale - np.float32(0.5)
        f = np.floor(src)
        lo = np.maximum(f.astype(np.int64), 0)
        hi = np.minimum(np.ceil(src).astype(np.int64), in_size - 1)
        return lo, hi, (src - f).astype(np.float32)

    ylo, yhi, yl = weights(out_h, h)
    xlo, xhi, xl = weights(out_w, w)
    tl, tr = img[:, ylo][:, :, xlo], img[:, ylo][:, :, xhi]
    bl, br = img[:, yhi][:, :, xlo], img[:, yhi][:, :, xhi]
    xl_ = xl[None, None, :, None]
    top = tl + (tr - tl) * xl_
    bot = bl + (br - bl) * xl_
    return (top + (bot - top) * yl[None, :, None, None]).astype(np.float32)


# ---- the reference's own logic -----------------------------------------------------------------------------------

def _one_hot_regions(cca, depth):
    """tf.one_hot(cca, depth, axis=0, dtype=bool)[1:]"""
    return np.stack([cca == i for i in range(1, depth)], 0) if depth > 1 else np.zeros((0,) + cca.shape, bool)


def separate_predictions(single_label, single_pred, thresholds, morph_filter_size=5):
    """metrics.py:108-162 ``_separate_predictions`` for one unbatched label / prediction pair."""
    thresholds = np.asarray(thresholds, np.float32).reshape(-1)
    nthr = len(thresholds)
    single_label = np.asarray(single_label) > 0.5
    cca_label = connected_components(single_label[None])[0]
    indiced_label = _one_hot_regions(cca_label, int(cca_label.max()) + 1)

    pred = np.broadcast_to(np.asarray(single_pred, np.float32), (nthr,) + np.shape(single_pred))
    pred = np.transpose(np.transpose(pred, (1, 2, 0)) >= thresholds, (2, 0, 1))
    pred = morph_open(pred.astype(np.int8)[..., None], morph_filter_size)[..., 0]
    cca_pred = connected_components(pred)                                   # [T,H,W], ids unique across thresholds
    # sparse.reduce_max over the non-zero entries of -cca_pred (0 where a threshold has no region at all)
    min_indices = np.array([-(np.max(-c[c != 0]) if np.any(c != 0) else 0) - 1 for c in cca_pred], np.int32)
    should_shift = (min_indices > 0).astype(np.int32)
    subtractor = (min_indices * should_shift)[:, None, None] * (cca_pred > 0).astype(np.int32)
    cca_pred = cca_pred - subtractor
    indiced_pred = _one_hot_regions(cca_pred, int(cca_pred.max()) + 1)      # [M,T,H,W]
    indiced_pred = np.transpose(indiced_pred, (0, 2, 3, 1))                 # [M,H,W,T]
    lengths = indiced_pred.any(axis=(1, 2))                                 # [M,T]
    existence_indicator = lengths.any(axis=0)
    if indiced_pred.shape[0] > 0:
        lengths = np.argmin(lengths.astype(np.uint8), axis=0)
        lengths = (lengths == 0) * existence_indicator * indiced_pred.shape[0] + lengths
    else:
        lengths = np.zeros(nthr, np.int64)
    return indiced_label, indiced_pred, lengths.astype(np.int64)


def iou_matrix(indiced_label, indiced_pred):
    """metrics.py:164-192 ``_IoU``: [N_label, M_pred, T] float32 (0/0 = nan cannot occur: label regions are non-empty)."""
    lab = indiced_label[:, None, :, :, None]
    prd = indiced_pred[None]
    inter = (lab & prd).sum(axis=(2, 3)).astype(np.float32)
    union = (lab | prd).sum(axis=(2, 3)).astype(np.float32)
    with np.errstate(invalid='ignore', divide='ignore'):
        return inter / union


def tp_fn_fp_single(single_label, single_pred, thresholds, iou_threshold=0.3, morph_filter_size=5):
    """metrics.py:275-288 ``_get_tp_fn_fp``."""
    indiced_label, indiced_pred, n_pred_masks = separate_predictions(single_label, single_pred, thresholds, morph_filter_size)
    iou = iou_matrix(indiced_label, indiced_pred)
    hit = iou > np.float32(iou_threshold)
    label_detected = hit.any(axis=1)                                        # [N,T]
    tp = label_detected.sum(axis=0).astype(np.int64)
    fn = (~label_detected).sum(axis=0).astype(np.int64)
    tp_pred = hit.any(axis=0).T                                             # [T,M]; ragged by n_pred_masks
    fp = np.array([int((~tp_pred[t, :n_pred_masks[t]]).sum()) for t in range(tp_pred.shape[0])], np.int64)
    tpp = np.array([int(tp_pred[t, :n_pred_masks[t]].sum()) for t in range(tp_pred.shape[0])], np.int64)
    return tp, fn, fp, tpp


def resize_pair(y_true, y_pred, resize_factor):
    """metrics.py:194-204, 210-214: stack (label, prediction), bilinear-resize both by ``resize_factor``."""
    y_true = np.asarray(y_true, np.float32)
    y_pred = np.asarray(y_pred, np.float32)
    if y_pred.ndim == 4:
        y_pred = y_pred[..., 0]
    both = np.stack([y_true, y_pred], axis=-1)                              # [B,W,H,2]
    tw, th = resize_target(both.shape[1], resize_factor), resize_target(both.shape[2], resize_factor)
    both = resize_bilinear(both, tw, th)
    return both[..., 0], both[..., 1]


def get_tp_fn_fp(y_true, y_pred, thresholds, iou_threshold=0.3, resize_factor=1.0, morph_filter_size=5, raw=False):
    """metrics.py:254-273 ``get_tp_fn_fp`` (sums over the batch); also returns the prediction-side TP count of
    ``get_tp_fp`` (metrics.py:229-252) as the fourth value."""
    lab, prd = resize_pair(y_true, y_pred, resize_factor)
    per = [tp_fn_fp_single(l, p, thresholds, iou_threshold, morph_filter_size) for l, p in zip(lab, prd)]
    nthr = len(np.asarray(thresholds).reshape(-1))
    arr = np.array(per, np.int64).reshape(len(per), 4, nthr)
    if raw:
        return arr
    s = arr.sum(axis=0)
    return s[0], s[1], s[2], s[3]
