"""Pixel-threshold metrics: host formulas against the oracle's brute-force restatement (CPU), and the
device histogram kernel against exact numpy counts (GPU, bit-exact)."""
import numpy as np
import pytest
import torch

from oracle import ref_metrics as rm


def _case(seed, n=20000, pos_rate=0.03):
    rng = np.random.default_rng(seed)
    y = (rng.uniform(size=n) < pos_rate).astype(np.float32)
    z = rng.normal(size=n) * 2 - 2 + 3 * y
    p = (1 / (1 + np.exp(-z))).astype(np.float32)
    # exact hits on thresholds and the extremes exercise the strict `>` comparison
    p[:7] = np.array([0.0, 1.0, 0.8, np.float32(1 / 149), np.float32(74 / 149), np.float32(148 / 149), 0.5], np.float32)
    return y, p


class _FakeCounts:
    """host-side stand-in for ThresholdCounts: same bins, filled by numpy"""

    def __init__(self, thresholds, y, p):
        thr = np.asarray(thresholds, np.float64).astype(np.float32)
        b = (thr[None, :] < p[:, None]).sum(1)
        n = len(thr)
        self.n = n
        self.h = np.concatenate([np.bincount(b[y != 0], minlength=n + 1), np.bincount(b[y == 0], minlength=n + 1)]).astype(np.int64)

    def counts(self):
        pos, neg = self.h[:self.n + 1], self.h[self.n + 1:]
        tp = pos[::-1].cumsum()[::-1][1:]
        fp = neg[::-1].cumsum()[::-1][1:]
        return tp, fp, pos.sum() - tp, neg.sum() - fp


@pytest.mark.parametrize('seed', [0, 1])
def test_host_metric_formulas_match_the_oracle(seed):
    from dnncancerannotator_b200.utils import metrics as M
    y, p = _case(seed)
    specs = [{'Precision': {'thresholds': 0.80, 'name': 'pixel/precision'}}, {'Recall': {'thresholds': 0.80, 'name': 'pixel/recall'}},
             {'AUC': {'curve': 'PR', 'name': 'pixel/AUPRC', 'num_thresholds': 150}},
             {'AUC': {'curve': 'ROC', 'name': 'pixel/AUROC', 'num_thresholds': 150}},
             {'FBetaScore': {'thresholds': 0.80, 'beta': 2.0, 'name': 'pixel/F2-score'}}]
    ms = [M.solve_metric(s) for s in specs]
    for m in ms:
        m._counts = _FakeCounts(m.thresholds, y, p)
        tp, fp, fn, tn = m._counts.counts()
        rtp, rfp, rfn, rtn = rm.confusion(y, p, m.thresholds)          # bins -> counts is exact
        assert np.array_equal(tp, rtp) and np.array_equal(fp, rfp) and np.array_equal(fn, rfn) and np.array_equal(tn, rtn)
    tp, fp, fn, tn = rm.confusion(y, p, [0.8])
    assert ms[0].result() == pytest.approx(float(rm.precision(tp, fp)[0]), rel=1e-12)
    assert ms[1].result() == pytest.approx(float(rm.recall(tp, fn)[0]), rel=1e-12)
    assert ms[4].result() == pytest.approx(float(rm.fbeta(tp, fp, fn, 2.0)[0]), rel=1e-12)
    thr = rm.auc_thresholds(150)
    assert np.allclose(thr, np.asarray(ms[2].thresholds))
    c = rm.confusion(y, p, thr)
    assert ms[2].result() == pytest.approx(rm.auc_pr(*c), rel=1e-12)
    assert ms[3].result() == pytest.approx(rm.auc_roc(*c), rel=1e-12)
    # definition-level checks: a perfect ranking has AUROC 1; precision/recall are plain ratios
    yy = np.array([0, 0, 1, 1], np.float32)
    pp = np.array([0.1, 0.2, 0.7, 0.9], np.float32)
    assert rm.auc_roc(*rm.confusion(yy, pp, thr)) == pytest.approx(1.0)
    tp, fp, fn, tn = rm.confusion(yy, np.array([0.9, 0.1, 0.9, 0.1], np.float32), [0.8])
    assert (tp[0], fp[0], fn[0], tn[0]) == (1, 1, 1, 1)
    assert isinstance(M.solve_metric({'RegionBasedRecall': {'thresholds': 0.8}}), M.RegionBasedRecall)


@pytest.mark.gpu
@pytest.mark.parametrize('n', [4 * 1024 * 37, 1000003])
def test_threshold_hist_kernel_bit_exact(n):
    from dnncancerannotator_b200.utils import metrics as M
    y, p = _case(7, n=n)
    yd, pd = torch.from_numpy(y).cuda(), torch.from_numpy(p).cuda()
    for thresholds in ([0.8], rm.auc_thresholds(150), [0.25, 0.5, 0.5, 0.75]):
        tc = M.ThresholdCounts(thresholds, 'cuda')
        tc.update(yd, pd)
        tc.update(yd, pd)                                  # accumulates
        torch.cuda.synchronize()
        tp, fp, fn, tn = tc.counts()
        rtp, rfp, rfn, rtn = rm.confusion(y, p, thresholds)
        assert np.array_equal(tp, 2 * rtp) and np.array_equal(fp, 2 * rfp) and np.array_equal(fn, 2 * rfn) and np.array_equal(tn, 2 * rtn)
        tc.reset()
        assert int(tc.hist.sum()) == 0


@pytest.mark.gpu
def test_model_evaluate_reports_compiled_metrics():
    """engine.py:273: the compiled metrics appear in evaluate()'s dict under their configured names"""
    from dnncancerannotator_b200.models import tf_models
    from dnncancerannotator_b200.synthetic import make_slices
    m = tf_models.UNetAnnotator(n_filters_first=3, n_downsample=2, rate=2, kernel_size=3, conv_stride=1, padding='same', dtype='bf16')
    m.build((None, 32, 32, 3))
    m.compile(metrics=[{'Precision': {'thresholds': 0.5, 'name': 'pixel/precision'}},
                       {'AUC': {'curve': 'ROC', 'num_thresholds': 50, 'name': 'pixel/AUROC'}},
                       {'RegionBasedRecall': {'thresholds': 0.8, 'name': 'region/recall'}}])
    x, y = make_slices(4, 32, 32, 3, seed=3)
    out = m.evaluate([(x, y)])
    probs = m._plan(4, 32, 32).probs.cpu().numpy().ravel()
    tp, fp, fn, tn = rm.confusion(y, probs, [0.5])
    assert out['pixel/precision'] == pytest.approx(float(rm.precision(tp, fp)[0]), abs=1e-12)
    assert out['pixel/AUROC'] == pytest.approx(rm.auc_roc(*rm.confusion(y, probs, rm.auc_thresholds(50))), abs=1e-12)
    from oracle import ref_region as rr
    pr = m._plan(4, 32, 32).probs.cpu().numpy()
    rtp, rfn, rfp, rtpp = rr.get_tp_fn_fp(y, pr, [0.8])
    assert out['region/recall'] == pytest.approx(float(rtp[0]) / (float(rtp[0] + rfn[0]) + 1e-7), rel=1e-6, abs=1e-7) and 'loss' in out
