"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the pixel-threshold metrics the reference
attaches to the model (engine.py:273, configs/additionals/metrics.yaml:1-23): tf.keras.metrics.Precision / Recall / AUC
(tensorflow==2.6, requirements.txt:2 -- not vendored, algorithm restated from its documented behaviour) and the
reference's own FBetaScore (annotator/utils/metrics.py:37-77).  parity unpinned: no TensorFlow here to generate vectors.

[TF-semantics] reproduced:
* confusion counts: y_true is cast to bool, a pixel is predicted positive iff y_pred > threshold (strict);
* AUC(num_thresholds=n): thresholds = [-1e-7] + [i/(n-1) for i in 1..n-2] + [1+1e-7], compared in float32;
  curve='ROC' with the default summation_method='interpolation' = trapezoids over (FPR, TPR);
  curve='PR' with 'interpolation' = Davis & Goadrich interpolation (keras `interpolate_pr_auc`);
* every ratio is a div_no_nan (0 when the denominator is 0).
Counts are kept exactly (int64) and the formulas evaluated in float64; Keras keeps float32 accumulators.
"""
import numpy as np

EPS = 1e-7


def auc_thresholds(num_thresholds=200):
    n = int(num_thresholds)
    return np.array([0.0 - EPS] + [(i + 1) * 1.0 / (n - 1) for i in range(n - 2)] + [1.0 + EPS], np.float64)


def confusion(y_true, y_pred, thresholds):
    """tp, fp, fn, tn per threshold by brute force (float32 comparisons)."""
    y = np.asarray(y_true).ravel() != 0
    p = np.asarray(y_pred, np.float32).ravel()
    thr = np.atleast_1d(np.asarray(thresholds, np.float64)).astype(np.float32)
    pos = p[None, :] > thr[:, None]
    tp = (pos & y[None, :]).sum(1).astype(np.int64)
    fp = (pos & ~y[None, :]).sum(1).astype(np.int64)
    return tp, fp, int(y.sum()) - tp, int((~y).sum()) - fp


def div_no_nan(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.where(b != 0, a / np.where(b != 0, b, 1.0), 0.0)


def precision(tp, fp):
    return div_no_nan(tp, tp + fp)


def recall(tp, fn):
    return div_no_nan(tp, tp + fn)


def fbeta(tp, fp, fn, beta, epsilon=1e-7):
    """annotator/utils/metrics.py:59-63"""
    p, r = precision(tp, fp), recall(tp, fn)
    return (1 + beta ** 2) * p * r / (beta ** 2 * p + r + epsilon)


def auc_roc(tp, fp, fn, tn):
    x = div_no_nan(fp, fp + tn)
    y = div_no_nan(tp, tp + fn)
    return float(np.sum((x[:-1] - x[1:]) * (y[:-1] + y[1:]) / 2.0))


def auc_pr(tp, fp, fn, tn):
    tp, fp, fn = (np.asarray(v, np.float64) for v in (tp, fp, fn))
    dtp = tp[:-1] - tp[1:]
    p = tp + fp
    dp = p[:-1] - p[1:]
    slope = div_no_nan(dtp, np.maximum(dp, 0))
    intercept = tp[1:] - slope * p[1:]
    ok = (p[:-1] > 0) & (p[1:] > 0)
    ratio = np.where(ok, div_no_nan(p[:-1], np.maximum(p[1:], 0)), 1.0)
    inc = div_no_nan(slope * (dtp + intercept * np.log(ratio)), np.maximum(tp[1:] + fn[1:], 0))
    return float(np.sum(inc))
