"""Pixel-threshold metrics of the reference on the device.

Mirrors ``annotator/utils/metrics.py:17-77`` (``solve_metric``, ``FBetaScore``) and the Keras metrics named by
``configs/additionals/metrics.yaml:1-23`` (``Precision``, ``Recall``, ``AUC``), which ``engine.py:273`` attaches to the
model.  Every one of them is a function of per-threshold confusion counts; the counts come from ONE device pass per
distinct threshold set (``dnnca_threshold_hist``), shared by all metrics that use that set.  The region-based metrics
(``metrics.py:80-194``: morphology + connected components) are out of scope and are skipped.
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


_REGISTRY = {'Precision': Precision, 'Recall': Recall, 'AUC': AUC, 'FBetaScore': FBetaScore}


def solve_metric(metric_spec):
    """``annotator/utils/metrics.py:17-33``: a one-entry dict {class_name: config} -> metric instance.  Region-based
    metrics (out of scope) give ``None``."""
    if isinstance(metric_spec, Metric):
        return metric_spec
    if isinstance(metric_spec, str):
        metric_spec = {metric_spec: {}}
    if not isinstance(metric_spec, dict) or len(metric_spec) != 1:
        raise ValueError(f'bad metric spec {metric_spec!r}')
    name, options = list(metric_spec.items())[0]
    if name.startswith('RegionBased'):
        warnings.warn(f'metric {name}: region-based metrics are not part of the device path, skipped')
        return None
    if name not in _REGISTRY:
        raise ValueError(f'unknown metric {name!r}')
    return _REGISTRY[name](**(options or {}))


class MetricSet:
    """The compiled metrics of a model: one ``ThresholdCounts`` per distinct threshold set."""

    def __init__(self, specs, device):
        self.metrics = [m for m in (solve_metric(s) for s in (specs or [])) if m is not None]
        self.groups = {}
        for m in self.metrics:
            if m.thresholds not in self.groups:
                self.groups[m.thresholds] = ThresholdCounts(m.thresholds, device)
            m._counts = self.groups[m.thresholds]

    def __bool__(self):
        return bool(self.metrics)

    def reset_state(self):
        for g in self.groups.values():
            g.reset()

    def update_state(self, y_true, y_pred):
        for g in self.groups.values():
            g.update(y_true, y_pred)

    def result(self):
        return {m.name: m.result() for m in self.metrics}
