"""Metrics of the reference on the device (SURVEY 8f N2 and the region-based "later" row).

Mirrors ``annotator/utils/metrics.py`` (``solve_metric`` :17-33, ``FBetaScore`` :37-77, the ``RegionBased*`` family
:80-520) and the Keras metrics named by ``configs/additionals/metrics.yaml`` (``Precision``, ``Recall``, ``AUC``), which
``engine.py:273`` attaches to the model.  The pixel metrics are functions of per-threshold confusion counts that come
from ONE device pass per distinct threshold set (``dnnca_threshold_hist``); the region metrics are functions of
per-threshold region detection counts that come from ONE device pass per distinct (thresholds, IoU threshold, resize
factor, opening size) configuration (``dnnca_region_confusion``: grey opening, union-find connected components, pair
histogram -- ``csrc/region_metrics.cu``), shared by all metrics that use it.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch

from .. import native as N

EPS = 1e-7


def _auc_thresholds(num_thresholds):
    n = int(num_thresholds)
    return tuple([0.0 - EPS] + [(i + 1) * 1.0 / (n - 1) for i in range(n - 2)] + [1.0 + EPS])


def _div_no_nan(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.where(b != 0, a / np.where(b != 0, b, 1.0), 0.0)


class ThresholdCounts:
    """Device histogram of predictions against one ascending threshold set (exact 64-bit counts)."""

    def __init__(self, thresholds, device):
        thr = np.asarray(thresholds, np.float64).astype(np.float32)
        assert thr.ndim == 1 and len(thr) >= 1 and np.all(np.diff(thr) >= 0), 'thresholds must be ascending'
        self.n = len(thr)
        self.thr = torch.from_numpy(thr).to(device)
        self.hist = torch.zeros(2 * (self.n + 1), dtype=torch.int64, device=device)

    def reset(self):
        self.hist.zero_()

    def update(self, y_true, y_pred):
        """y_true / y_pred: float32 device tensors of equal element count (labels, probabilities)."""
        assert y_true.numel() == y_pred.numel() and y_true.dtype == torch.float32 and y_pred.dtype == torch.float32
        y_true, y_pred = y_true.contiguous(), y_pred.contiguous()
        N.call('dnnca_threshold_hist', N.stream_ptr(), N.ptr(y_pred), N.ptr(y_true), y_pred.numel(), N.ptr(self.thr), self.n,
               N.ptr(self.hist))

    def counts(self):
        """(tp, fp, fn, tn) per threshold as int64 numpy arrays."""
        h = self.hist.cpu().numpy()
        pos, neg = h[:self.n + 1], h[self.n + 1:]
        # bin b = number of thresholds strictly below p, so p > t_k  <=>  b > k
        tp = pos[::-1].cumsum()[::-1][1:]
        fp = neg[::-1].cumsum()[::-1][1:]
        return tp, fp, pos.sum() - tp, neg.sum() - fp


class Metric:
    thresholds: tuple = ()

    def __init__(self, name):
        self.name = name
        self._counts = None            # ThresholdCounts, attached by MetricSet

    def result(self):
        raise NotImplementedError

    def get_config(self):
        return {'name': self.name}


def _thr_tuple(thresholds):
    t = np.atleast_1d(np.asarray(thresholds, np.float64))
    return tuple(float(v) for v in t)


class Precision(Metric):
    """tf.keras.metrics.Precision(thresholds=...)"""

    def __init__(self, thresholds=0.5, name='precision', **kargs):
        super().__init__(name)
        self.thresholds = _thr_tuple(thresholds)

    def result(self):
        tp, fp, fn, tn = self._counts.counts()
        r = _div_no_nan(tp, tp + fp)
        return float(r[0]) if len(r) == 1 else r


class Recall(Metric):
    """tf.keras.metrics.Recall(thresholds=...)"""

    def __init__(self, thresholds=0.5, name='recall', **kargs):
        super().__init__(name)
        self.thresholds = _thr_tuple(thresholds)

    def result(self):
        tp, fp, fn, tn = self._counts.counts()
        r = _div_no_nan(tp, tp + fn)
        return float(r[0]) if len(r) == 1 else r


class FBetaScore(Metric):
    """annotator/utils/metrics.py:37-77"""

    def __init__(self, beta, thresholds, epsilon=1e-07, name='fbeta', **kargs):
        super().__init__(name)
        assert beta > 0
        self.beta, self.epsilon = beta, epsilon
        self.thresholds = _thr_tuple(thresholds)

    def result(self):
        tp, fp, fn, tn = self._counts.counts()
        p, r = _div_no_nan(tp, tp + fp), _div_no_nan(tp, tp + fn)
        s = (1 + self.beta ** 2) * p * r / (self.beta ** 2 * p + r + self.epsilon)
        return float(s[0]) if len(s) == 1 else s


class AUC(Metric):
    """tf.keras.metrics.AUC(num_thresholds, curve='ROC'|'PR', summation_method='interpolation')"""

    def __init__(self, num_thresholds=200, curve='ROC', summation_method='interpolation', name='auc', **kargs):
        super().__init__(name)
        if summation_method != 'interpolation':
            raise NotImplementedError(f'AUC summation_method {summation_method!r} (the reference uses the default)')
        if str(curve).upper() not in ('ROC', 'PR'):
            raise ValueError(f'AUC curve {curve!r}')
        self.curve = str(curve).upper()
        self.num_thresholds = int(num_thresholds)
        self.thresholds = _auc_thresholds(num_thresholds)

    def result(self):
        tp, fp, fn, tn = (v.astype(np.float64) for v in self._counts.counts())
        if self.curve == 'ROC':
            x, y = _div_no_nan(fp, fp + tn), _div_no_nan(tp, tp + fn)
            return float(np.sum((x[:-1] - x[1:]) * (y[:-1] + y[1:]) / 2.0))
        dtp = tp[:-1] - tp[1:]
        p = tp + fp
        slope = _div_no_nan(dtp, np.maximum(p[:-1] - p[1:], 0))
        intercept = tp[1:] - slope * p[1:]
        ok = (p[:-1] > 0) & (p[1:] > 0)
        ratio = np.where(ok, _div_no_nan(p[:-1], np.maximum(p[1:], 0)), 1.0)
        return float(np.sum(_div_no_nan(slope * (dtp + intercept * np.log(ratio)), np.maximum(tp[1:] + fn[1:], 0))))


# ---- region-based metrics (annotator/utils/metrics.py:80-520) -------------------------------------------------------

def _resize_target(size, resize_factor):
    """metrics.py:199-200: tf.cast(tf.cast(size, tf.float16) * resize_factor, tf.int32)"""
    return int(np.float16(np.float16(size) * np.float16(resize_factor)))


class RegionCounts:
    """Device counters behind every region-based metric of one configuration.  ``totals`` is int64 [4, T] in the
    order of ``dnnca_region_confusion``: labels detected, labels missed, predictions without a hit, predictions with a
    hit."""

    MAX_WORKSPACE = 1 << 30

    def __init__(self, thresholds, iou_threshold, resize_factor, morph_filter_size, device, table_slots=4096):
        thr = np.asarray(thresholds, np.float64).reshape(-1).astype(np.float32)
        assert len(thr) >= 1 and np.all(thr >= 0), 'thresholds must be non-negative (metrics.py:94)'
        self.n = len(thr)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise N.DnncaError('region-based metrics run on the CUDA device only (no CPU path exists)')
        self.thr = torch.from_numpy(thr).to(self.device)
        self.iou_threshold = float(iou_threshold)
        self.resize_factor = float(resize_factor)
        self.morph = int(morph_filter_size)
        self.slots = int(table_slots)
        self.totals = torch.zeros(4, self.n, dtype=torch.int64, device=self.device)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._ws = None

    def reset(self):
        self.totals.zero_()
        self.overflow.zero_()

    def _f32(self, t, squeeze_last):
        t = torch.as_tensor(t)
        if squeeze_last and t.dim() == 4:
            assert t.shape[-1] == 1, 'predictions are [B,H,W,1] (metrics.py:209)'
            t = t[..., 0]
        assert t.dim() == 3, 'labels / predictions are batches of 2-D images'
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    def _resized(self, t):
        if self.resize_factor == 1.0:            # tf.image.resize to the same size is the identity
            return t
        n, h, w = t.shape
        oh, ow = _resize_target(h, self.resize_factor), _resize_target(w, self.resize_factor)
        assert oh > 0 and ow > 0, 'resize_factor leaves no pixels'
        out = torch.empty(n, oh, ow, dtype=torch.float32, device=self.device)
        N.call('dnnca_resize_bilinear', N.stream_ptr(), N.ptr(t), n, h, w, N.ptr(out), oh, ow)
        return out

    def update(self, y_true, y_pred, accumulate=True, raw=False):
        """One pass over a batch.  Returns the per-slice counters int32 [B, 4, T] when ``raw`` (device tensor)."""
        lab, prd = self._resized(self._f32(y_true, False)), self._resized(self._f32(y_pred, True))
        assert lab.shape == prd.shape, f'label {tuple(lab.shape)} and prediction {tuple(prd.shape)} shapes differ'
        n, h, w = lab.shape
        lib = N.lib()
        per1 = int(lib.dnnca_region_workspace_bytes(1, h, w, self.n, self.slots))
        chunk = max(1, min(n, self.MAX_WORKSPACE // max(per1, 1), 65535 // self.n))
        need = int(lib.dnnca_region_workspace_bytes(chunk, h, w, self.n, self.slots))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        per_slice = torch.zeros(n, 4, self.n, dtype=torch.int32, device=self.device) if raw else None
        for b0 in range(0, n, chunk):
            m = min(chunk, n - b0)
            N.call('dnnca_region_confusion', N.stream_ptr(), N.ptr(lab[b0:b0 + m]), N.ptr(prd[b0:b0 + m]), m, h, w, N.ptr(self.thr),
                   self.n, self.iou_threshold, self.morph, N.ptr(self._ws), self._ws.numel(), self.slots,
                   N.ptr(per_slice[b0:b0 + m]) if raw else None, N.ptr(self.totals) if accumulate else None,
                   N.ptr(self.overflow))
        return per_slice

    def check(self):
        if int(self.overflow.item()):
            raise N.DnncaError(f'region metrics: more than {self.slots} distinct overlapping (label region, prediction '
                               'region) pairs in one slice; construct the metric with a larger table_slots')

    def counts(self):
        """(tp, fn, fp, tp_pred) int64 numpy arrays of length T."""
        t = self.totals.cpu().numpy()
        self.check()
        return t[0], t[1], t[2], t[3]


class _RegionBasedMetric(Metric):
    """metrics.py:80-303: common part of the region-based metrics.

    Args (as in the reference): ``thresholds`` scalar or vector; ``IoU_threshold`` minimum IoU between a prediction
    region and a label region to count as a detection; ``epsilon``; ``resize_factor`` bilinear shrink of both images
    before the analysis; ``morph_filter_size`` side of the square opening applied to the thresholded prediction.
    ``table_slots`` (this build only) bounds the distinct overlapping region pairs per slice and threshold.
    """

    def __init__(self, thresholds, IoU_threshold=0.30, epsilon=1e-07, resize_factor=1.0, morph_filter_size=5, name=None,
                 table_slots=4096, **kargs):
        super().__init__(name or type(self).__name__)
        thr = _thr_tuple(thresholds)
        assert all(t >= 0 for t in thr), 'thresholds must be non-negative (metrics.py:94)'
        self.thresholds = thr
        self.IoU_threshold, self.epsilon = IoU_threshold, epsilon
        self.resize_factor, self.morph_filter_size, self.table_slots = resize_factor, morph_filter_size, table_slots

    # shared-engine key: metrics of equal configuration use one device pass
    @property
    def region_key(self):
        return (self.thresholds, float(self.IoU_threshold), float(self.resize_factor), int(self.morph_filter_size),
                int(self.table_slots))

    def _engine(self, like=None):
        """The device counters: attached by ``MetricSet`` or created on first use for a stand-alone metric."""
        if self._counts is None:
            if torch.is_tensor(like) and like.is_cuda:
                dev = like.device
            elif torch.cuda.is_available():
                dev = torch.device('cuda', torch.cuda.current_device())
            else:
                raise N.DnncaError('region-based metrics need a CUDA device (no CPU path exists)')
            self._counts = RegionCounts(self.thresholds, self.IoU_threshold, self.resize_factor, self.morph_filter_size, dev,
                                        self.table_slots)
        return self._counts

    def _batch(self, y_true, y_pred, sample_weight):
        if sample_weight is not None:
            raise NotImplementedError                      # metrics.py:208
        eng = self._engine(y_pred)
        raw = eng.update(y_true, y_pred, accumulate=False, raw=True).cpu().numpy().astype(np.int64)
        eng.check()
        return raw                                          # [B, 4, T]

    def get_tp_fn(self, y_true, y_pred, sample_weight=None):
        """metrics.py:206-227: (label regions detected, label regions missed) per threshold, summed over the batch."""
        r = self._batch(y_true, y_pred, sample_weight).sum(0)
        return r[0], r[1]

    def get_tp_fp(self, y_true, y_pred, sample_weight=None):
        """metrics.py:229-252: (prediction regions that hit a label, prediction regions that hit none)."""
        r = self._batch(y_true, y_pred, sample_weight).sum(0)
        return r[3], r[2]

    def get_tp_fn_fp(self, y_true, y_pred, sample_weight=None, return_raw=False):
        """metrics.py:254-288: (labels detected, labels missed, predictions without a hit); per slice if ``return_raw``."""
        r = self._batch(y_true, y_pred, sample_weight)
        if return_raw:
            return r[:, 0], r[:, 1], r[:, 2]
        r = r.sum(0)
        return r[0], r[1], r[2]

    def update_state(self, y_true, y_pred, sample_weight=None):
        if sample_weight is not None:
            raise NotImplementedError
        self._engine(y_pred).update(y_true, y_pred, accumulate=True, raw=False)

    def reset_state(self):
        if self._counts is not None:
            self._counts.reset()

    def _ratio(self, num, other):
        r = num.astype(np.float32) / ((num + other).astype(np.float32) + np.float32(self.epsilon))
        return float(r[0]) if len(r) == 1 else r

    @staticmethod
    def _squeeze(v):
        return int(v[0]) if len(v) == 1 else v

    def get_config(self):
        """metrics.py:290-296 (+ the opening size, which the reference's get_config drops)."""
        c = super().get_config()
        thr = self.thresholds
        c.update(thresholds=thr[0] if len(thr) == 1 else list(thr), IoU_threshold=self.IoU_threshold, epsilon=self.epsilon,
                 resize_factor=self.resize_factor, morph_filter_size=self.morph_filter_size)
        return c


class RegionBasedRecall(_RegionBasedMetric):
    """metrics.py:336-362"""

    def result(self):
        tp, fn, fp, tpp = self._engine().counts()
        return self._ratio(tp, fn)


class RegionBasedPrecision(_RegionBasedMetric):
    """metrics.py:365-391 (true positives counted on the prediction side, get_tp_fp)"""

    def result(self):
        tp, fn, fp, tpp = self._engine().counts()
        return self._ratio(tpp, fp)


class RegionBasedTruePositives(_RegionBasedMetric):
    """metrics.py:394-414"""

    def result(self):
        return self._squeeze(self._engine().counts()[0])


class RegionBasedFalsePositives(_RegionBasedMetric):
    """metrics.py:420-440"""

    def result(self):
        return self._squeeze(self._engine().counts()[2])


class RegionBasedFalseNegatives(_RegionBasedMetric):
    """metrics.py:443-463"""

    def result(self):
        return self._squeeze(self._engine().counts()[1])


class RegionBasedFBetaScore(_RegionBasedMetric):
    """metrics.py:306-333: FBetaScore over RegionBasedPrecision / RegionBasedRecall."""

    def __init__(self, beta, thresholds, IoU_threshold=0.30, epsilon=1e-07, resize_factor=1.0, **kargs):
        super().__init__(thresholds, IoU_threshold, epsilon, resize_factor, **kargs)
        assert beta > 0
        self.beta = beta

    def result(self):
        tp, fn, fp, tpp = self._engine().counts()
        eps = np.float32(self.epsilon)
        p = tpp.astype(np.float32) / ((tpp + fp).astype(np.float32) + eps)
        r = tp.astype(np.float32) / ((tp + fn).astype(np.float32) + eps)
        b2 = np.float32(self.beta ** 2)
        s = (1 + b2) * p * r / (b2 * p + r + eps)
        return float(s[0]) if len(s) == 1 else s

    def get_config(self):
        c = super().get_config()
        c['beta'] = self.beta
        return c


class RegionBasedConfusionMatrix(_RegionBasedMetric):
    """metrics.py:466-512: all three counters of ``get_tp_fn_fp``; ``result_dict`` is what the Visualizer reads."""

    def result(self):
        return float('nan')

    def result_dict(self):
        tp, fn, fp, tpp = self._engine().counts()
        return {'true_positive_counts': self._squeeze(tp), 'false_positive_counts': self._squeeze(fp),
                'false_negative_counts': self._squeeze(fn), 'recall': self._ratio(tp, fn), 'precision': self._ratio(tp, fp)}


_REGISTRY = {'Precision': Precision, 'Recall': Recall, 'AUC': AUC, 'FBetaScore': FBetaScore,
             'RegionBasedRecall': RegionBasedRecall, 'RegionBasedPrecision': RegionBasedPrecision,
             'RegionBasedFBetaScore': RegionBasedFBetaScore, 'RegionBasedTruePositives': RegionBasedTruePositives,
             'RegionBasedFalsePositives': RegionBasedFalsePositives, 'RegionBasedFalseNegatives': RegionBasedFalseNegatives,
             'RegionBasedConfusionMatrix': RegionBasedConfusionMatrix}


def solve_metric(metric_spec):
    """``annotator/utils/metrics.py:17-33``: a one-entry dict {class_name: config} -> metric instance."""
    if isinstance(metric_spec, Metric):
        return metric_spec
    if isinstance(metric_spec, str):
        metric_spec = {metric_spec: {}}
    if not isinstance(metric_spec, dict) or len(metric_spec) != 1:
        raise ValueError(f'bad metric spec {metric_spec!r}')
    name, options = list(metric_spec.items())[0]
    if name not in _REGISTRY:
        raise ValueError(f'unknown metric {name!r}')
    return _REGISTRY[name](**(options or {}))


class MetricSet:
    """The compiled metrics of a model: one ``ThresholdCounts`` per distinct threshold set."""

    def __init__(self, specs, device):
        self.metrics = [m for m in (solve_metric(s) for s in (specs or [])) if m is not None]
        self.groups = {}
        self.region_groups = {}
        for m in self.metrics:
            if isinstance(m, _RegionBasedMetric):
                if m.region_key not in self.region_groups:
                    self.region_groups[m.region_key] = RegionCounts(m.thresholds, m.IoU_threshold, m.resize_factor,
                                                                    m.morph_filter_size, device, m.table_slots)
                m._counts = self.region_groups[m.region_key]
                continue
            if m.thresholds not in self.groups:
                self.groups[m.thresholds] = ThresholdCounts(m.thresholds, device)
            m._counts = self.groups[m.thresholds]

    def __bool__(self):
        return bool(self.metrics)

    def reset_state(self):
        for g in list(self.groups.values()) + list(self.region_groups.values()):
            g.reset()

    def update_state(self, y_true, y_pred):
        for g in self.groups.values():
            g.update(y_true, y_pred)
        for g in self.region_groups.values():          # y_true [B,H,W], y_pred [B,H,W,1]
            g.update(y_true, y_pred)

    def result(self):
        return {m.name: m.result() for m in self.metrics}
