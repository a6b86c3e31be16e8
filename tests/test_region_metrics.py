"""Region-based metrics (SURVEY 8f "later" row): the reference's own test module
``annotator/tests/test_region_metrics.py`` restated test for test (same sample generators, same expected counts -- they
are known by construction), run against

* the numpy oracle ``oracle/ref_region.py`` (CPU, pins the oracle: the only path of this repo for which the reference
  ships known-answer tests), and
* the CUDA path ``dnncancerannotator_b200.utils.metrics.RegionBased*`` -> ``dnnca_region_confusion`` (``-m gpu``),

plus cross-checks of the restated third-party ops against independent implementations (scipy / torch) and bit-exact
parity of the CUDA path with the oracle on seeded random slices, edge cases included (empty batch members, regions on the
image border, image sizes that are not multiples of the warp size, overflow of the pair table).
"""
import random
from copy import deepcopy

import numpy as np
import pytest
import torch

from oracle import ref_region as rr


# ---- the sample generators of the reference's tests (test_region_metrics.py:274-372), numpy instead of tf ----------

def draw_circle(tensor, radius, center_x, center_y, min_=1.0, max_=1.0, rnd=random):
    """test_region_metrics.py:351-372"""
    assert tensor.ndim == 2
    width, height = tensor.shape
    dt = tensor.dtype
    center_x, center_y = np.asarray(center_x).astype(dt), np.asarray(center_y).astype(dt)
    x_dist = (np.arange(width, dtype=dt) - center_x) ** 2
    x_dist = np.broadcast_to(x_dist, (width, width))
    y_dist = (np.arange(height, dtype=dt) - center_y) ** 2
    y_dist = np.broadcast_to(y_dist, (height, height)).T
    dist = np.sqrt((x_dist + y_dist).astype(np.float32))
    output = (dist < np.float32(radius)).astype(dt)
    output = (output.astype(np.float32) * np.float32(rnd.uniform(min_, max_))).astype(dt)
    return output + tensor


class Samples:
    """setUp + generate_* of TestRegionMetricsSingleThreshold (test_region_metrics.py:19-35, 274-349)"""

    def __init__(self, seed, batch_size=10, size=200):
        self.rng = np.random.default_rng(seed)
        self.rnd = random.Random(seed)
        self.batch_size = batch_size
        self.radius = self.rng.integers(10, 30, batch_size)
        self.center_x = self.rng.integers(30, 70, batch_size)
        self.center_y = self.rng.integers(80, 120, batch_size)
        self.center_x_off = self.rng.integers(130, 170, batch_size)
        self.center_y_off = self.rng.integers(80, 120, batch_size)
        self.width = self.height = size

    def _circles(self, cxs, cys):
        return np.stack([draw_circle(np.zeros((self.width, self.height), np.int64), r, cx, cy, rnd=self.rnd)
                         for r, cx, cy in zip(self.radius, cxs, cys)], 0)

    def _indicator(self, n_one, dtype):
        ind = np.concatenate([np.ones(n_one, dtype), np.zeros(self.batch_size - n_one, dtype)])
        self.rng.shuffle(ind)
        return ind

    def tp_fn(self, tp_rate):
        y_true = self._circles(self.center_x, self.center_y)
        y_pred = y_true.astype(np.float32)[..., None]
        n_tp = int(self.batch_size * tp_rate)
        y_pred = y_pred * self._indicator(n_tp, np.float32)[:, None, None, None]
        return y_true, y_pred, n_tp, self.batch_size - n_tp

    def tp_fp(self, tp_rate):
        y_true = self._circles(self.center_x, self.center_y)
        y_pred = y_true.astype(np.float32)[..., None]
        n_tp = int(self.batch_size * tp_rate)
        y_true = y_true * self._indicator(n_tp, y_true.dtype)[:, None, None]
        return y_true, y_pred, n_tp, self.batch_size - n_tp

    def off(self, off_rate):
        offs = self._circles(self.center_x_off, self.center_y_off).astype(np.float32)[..., None]
        n_off = int(self.batch_size * off_rate)
        return offs * self._indicator(n_off, np.float32)[:, None, None, None], n_off

    def null(self):
        y_true = np.zeros((self.batch_size, self.width, self.height), np.int64)
        return y_true, y_true.astype(np.float32)[..., None]

    def random(self, nslices, min_=1.0, max_=1.0):
        def gen_slice(dtype, ncircles, lo=1.0, hi=1.0):
            image = np.zeros((self.width, self.height), dtype)
            for _ in range(ncircles):
                image = draw_circle(image, self.rnd.uniform(5.0, self.width / 20), self.rnd.uniform(0.0, self.width),
                                    self.rnd.uniform(0.0, self.height), lo, hi, rnd=self.rnd)
            return image
        y_true = np.stack([gen_slice(np.int32, 5) for _ in range(nslices)], 0)
        y_pred = np.stack([gen_slice(np.float32, 5, min_, max_) for _ in range(nslices)], 0)
        return y_true, y_pred[..., None]


# ---- two interchangeable back ends with the reference metric's method names ---------------------------------------------

class OracleMetric:
    def __init__(self, thresholds, IoU_threshold=0.3, resize_factor=1.0, **kargs):
        self.cfg = dict(thresholds=thresholds, IoU_threshold=IoU_threshold, resize_factor=resize_factor)
        self.thr = np.atleast_1d(np.asarray(thresholds, np.float64))
        self.tot = np.zeros((4, len(self.thr)), np.int64)

    def _all(self, y_true, y_pred):
        return rr.get_tp_fn_fp(y_true, y_pred, self.thr, self.cfg['IoU_threshold'], self.cfg['resize_factor'])

    def get_tp_fn(self, y_true, y_pred, sw=None):
        tp, fn, fp, tpp = self._all(y_true, y_pred)
        return tp, fn

    def get_tp_fp(self, y_true, y_pred, sw=None):
        tp, fn, fp, tpp = self._all(y_true, y_pred)
        return tpp, fp

    def get_tp_fn_fp(self, y_true, y_pred, sw=None):
        tp, fn, fp, tpp = self._all(y_true, y_pred)
        return tp, fn, fp

    def get_config(self):
        return dict(self.cfg)


def make_metric(backend, **cfg):
    if backend == 'oracle':
        return OracleMetric(**cfg)
    from dnncancerannotator_b200.utils import metrics as M
    return M.RegionBasedConfusionMatrix(**cfg)


BACKENDS = ['oracle', 'cuda']
# (n_threshold, resize_factor): TestRegionMetrics{Single,Multi}Threshold[Shrinked] (test_region_metrics.py:18, 375, 400, 408)
VARIANTS = [(1, 1.0), (10, 1.0), (1, 0.5), (10, 0.5)]


def thresholds_of(n_threshold):
    if n_threshold == 1:
        return 0.5
    thr = [i / (n_threshold - 1) for i in range(n_threshold)]
    thr[0] = 0.001
    return thr


@pytest.fixture(params=[pytest.param((b, v), marks=[pytest.mark.gpu] if b == 'cuda' else [], id=f'{b}-T{v[0]}-r{v[1]}')
                        for b in BACKENDS for v in VARIANTS])
def setup(request):
    backend, (n_thr, resize) = request.param
    # the oracle materialises [labels, preds, H, W, T] like the reference: a smaller canvas keeps the CPU suite short
    size, batch = (200, 10) if backend == 'cuda' else (100, 6)
    s = Samples(seed=11, batch_size=batch, size=size)
    if size != 200:                 # same construction, scaled to the canvas
        s.radius, s.center_x, s.center_y = s.radius // 2 + 4, s.center_x // 2, s.center_y // 2
        s.center_x_off, s.center_y_off = s.center_x_off // 2, s.center_y_off // 2
    metric = make_metric(backend, thresholds=thresholds_of(n_thr), IoU_threshold=0.3, resize_factor=resize)
    return backend, s, metric, n_thr


def L(v):
    return np.asarray(v).reshape(-1).tolist()


@pytest.mark.parametrize('rate', [1.0, 0.0, 0.5])
def test_tp_fn(setup, rate):
    """test_tp_fn_all_tp / all_fn / half (test_region_metrics.py:43-62, 71-77)"""
    _, s, metric, T = setup
    y_true, y_pred, n_tp, n_fn = s.tp_fn(rate)
    tp, fn = metric.get_tp_fn(y_true, y_pred, None)
    assert L(tp) == [n_tp] * T and L(fn) == [n_fn] * T


def test_tp_fn_all_fp(setup):
    """test_region_metrics.py:57-62: no label regions -> nothing to detect"""
    _, s, metric, T = setup
    y_true, y_pred, _, _ = s.tp_fp(0.0)
    tp, fn = metric.get_tp_fn(y_true, y_pred, None)
    assert L(tp) == [0] * T and L(fn) == [0] * T


@pytest.mark.parametrize('rate', [0.0, 1.0, 0.5])
def test_tp_fp(setup, rate):
    """test_tp_fp_all_tp / all_fp / half (test_region_metrics.py:71-98).  NB the reference names the generator's
    argument tp_rate but n_tp = batch * rate are the slices that KEEP their label."""
    _, s, metric, T = setup
    y_true, y_pred, n_tp, n_fp = s.tp_fp(rate)
    tp, fp = metric.get_tp_fp(y_true, y_pred, None)
    assert L(tp) == [n_tp] * T and L(fp) == [n_fp] * T


def test_tp_fp_all_fn(setup):
    """test_region_metrics.py:85-90"""
    _, s, metric, T = setup
    y_true, y_pred, _, _ = s.tp_fn(0.0)
    tp, fp = metric.get_tp_fp(y_true, y_pred, None)
    assert L(tp) == [0] * T and L(fp) == [0] * T


@pytest.mark.parametrize('rate', [0.0, 1.0, 0.5])
def test_tp_fn_fp(setup, rate):
    """test_tp_fn_fp_all_tp / all_fp / half (test_region_metrics.py:100-130)"""
    _, s, metric, T = setup
    y_true, y_pred, n_tp, n_fp = s.tp_fp(rate)
    tp, fn, fp = metric.get_tp_fn_fp(y_true, y_pred, None)
    assert L(tp) == [n_tp] * T and L(fn) == [0] * T and L(fp) == [n_fp] * T


def test_tp_fn_fp_all_fn(setup):
    """test_region_metrics.py:116-122"""
    _, s, metric, T = setup
    y_true, y_pred, n_tp, n_fn = s.tp_fn(0.0)
    tp, fn, fp = metric.get_tp_fn_fp(y_true, y_pred, None)
    assert L(tp) == [0] * T and L(fn) == [n_fn] * T and L(fp) == [0] * T


def test_tp_fn_fp_null(setup):
    """test_region_metrics.py:132-138"""
    _, s, metric, T = setup
    y_true, y_pred = s.null()
    tp, fn, fp = metric.get_tp_fn_fp(y_true, y_pred, None)
    assert L(tp) == [0] * T and L(fn) == [0] * T and L(fp) == [0] * T


def test_tp_fn_fp_mixed(setup):
    """test_region_metrics.py:140-149"""
    _, s, metric, T = setup
    y_true, y_pred, n_tp, n_fn = s.tp_fn(0.4)
    offs, n_off = s.off(0.7)
    tp, fn, fp = metric.get_tp_fn_fp(y_true, y_pred + offs, None)
    assert L(tp) == [n_tp] * T and L(fn) == [n_fn] * T and L(fp) == [n_off] * T


def test_consistency(setup):
    """test_consistency / test_consistency_random (test_region_metrics.py:151-175)"""
    backend, s, metric, T = setup
    y_true, y_pred, _, _ = s.tp_fn(0.4)
    offs, _ = s.off(0.7)
    cases = [(y_true, y_pred + offs)] + [s.random(20 if backend == 'cuda' else 3) for _ in range(10 if backend == 'cuda' else 1)]
    for yt, yp in cases:
        tp, fn, fp = metric.get_tp_fn_fp(yt, yp, None)
        tp2, fn2 = metric.get_tp_fn(yt, yp, None)
        _, fp2 = metric.get_tp_fp(yt, yp, None)
        assert L(tp) == L(tp2) and L(fn) == L(fn2) and L(fp) == L(fp2)


def test_consistency_multithresholds(setup):
    """test_region_metrics.py:390-403: one metric over T thresholds == T single-threshold metrics"""
    backend, s, metric, T = setup
    if T == 1:
        pytest.skip('multi-threshold variant only')
    y_true, y_pred = s.random(20 if backend == 'cuda' else 3, 0.2, 1.0)
    tp, fn, fp = metric.get_tp_fn_fp(y_true, y_pred, None)
    cfg = metric.get_config()
    singles = []
    for t in np.atleast_1d(cfg['thresholds']):
        c = deepcopy(cfg)
        c['thresholds'] = [float(t)]
        singles.append(make_metric(backend, **c).get_tp_fn_fp(y_true, y_pred, None))
    assert [int(v[0][0]) for v in singles] == L(tp)
    assert [int(v[1][0]) for v in singles] == L(fn)
    assert [int(v[2][0]) for v in singles] == L(fp)


# ---- the restated third-party ops against independent implementations (pins the oracle's [TF-semantics] pieces) -----

def _blobs(seed, n, h, w, density=0.55):
    rng = np.random.default_rng(seed)
    from scipy import ndimage
    return ndimage.uniform_filter(rng.uniform(size=(n, h, w)), size=(1, 3, 3)) > density


def test_oracle_connected_components_matches_scipy():
    from scipy import ndimage
    masks = _blobs(0, 3, 37, 45, 0.52)
    ids = rr.connected_components(masks)
    base = 0
    for b in range(len(masks)):
        ref, n = ndimage.label(masks[b])                    # default structure: the 4-neighbourhood cross
        assert np.array_equal(ids[b], np.where(ref > 0, ref + base, 0))     # raster order of first pixels, ids continue
        base += n
    # known answer: diagonal neighbours are separate components under 4-connectivity, a row run is one
    k = np.zeros((1, 3, 4), bool)
    k[0, 0, 0] = k[0, 1, 1] = True
    k[0, 2, 2:4] = True
    assert rr.connected_components(k)[0].tolist() == [[1, 0, 0, 0], [0, 2, 0, 0], [0, 0, 3, 3]]


@pytest.mark.parametrize('k', [3, 5, 4])
def test_oracle_morph_open_matches_scipy(k):
    from scipy import ndimage
    x = _blobs(1, 2, 33, 29, 0.5).astype(np.int8)
    got = rr.morph_open(x[..., None], k)[..., 0]
    origin = 0 if k % 2 else -1                              # even windows: TF pads (k-1)//2 before
    for b in range(len(x)):
        er = ndimage.minimum_filter(x[b], size=k, mode='constant', cval=1, origin=origin)
        di = ndimage.maximum_filter(er, size=k, mode='constant', cval=0, origin=origin)
        assert np.array_equal(got[b], di)
    # opening keeps a k x k square and removes anything thinner; a square touching the border is not eroded from outside
    z = np.zeros((1, 12, 12), np.int8)
    z[0, 0:5, 0:5] = 1
    z[0, 8, :] = 1
    o = rr.morph_open(z, 5)
    assert o[0, 0:5, 0:5].all() and o[0, 8].sum() == 0


def test_oracle_resize_matches_torch_and_box_average():
    rng = np.random.default_rng(2)
    x = rng.uniform(size=(2, 20, 26, 2)).astype(np.float32)
    for oh, ow in [(10, 13), (7, 9), (20, 26), (31, 40)]:
        ref = torch.nn.functional.interpolate(torch.from_numpy(x).permute(0, 3, 1, 2), size=(oh, ow), mode='bilinear',
                                              align_corners=False).permute(0, 2, 3, 1).numpy()
        assert np.allclose(rr.resize_bilinear(x, oh, ow), ref, atol=2e-6)
    half = rr.resize_bilinear(x, 10, 13)                    # factor 0.5 = exact 2x2 box average
    box = x.reshape(2, 10, 2, 13, 2, 2).mean(axis=(2, 4))
    assert np.allclose(half, box, atol=1e-7)
    assert np.array_equal(rr.resize_bilinear(x, 20, 26), x)
    assert rr.resize_target(200, 0.5) == 100 and rr.resize_target(256, 0.5) == 128 and rr.resize_target(101, 0.5) == 50


def test_oracle_iou_known_answer():
    """two 4x4 squares overlapping in a 2x4 strip: IoU = 8 / 24; one label, one hit at 0.3, none at 0.34"""
    lab = np.zeros((1, 16, 16), np.float32)
    prd = np.zeros((1, 16, 16, 1), np.float32)
    lab[0, 2:6, 2:6] = 1
    prd[0, 4:8, 2:6, 0] = 1
    for iou_thr, want in [(0.3, (1, 0, 0, 1)), (0.34, (0, 1, 1, 0))]:
        tp, fn, fp, tpp = rr.get_tp_fn_fp(lab, prd, [0.5], iou_thr, morph_filter_size=3)
        assert (tp[0], fn[0], fp[0], tpp[0]) == want


# ---- CUDA path against the oracle, bit-exact ---------------------------------------------------------------------------

def _random_case(seed, n, h, w):
    """blobby labels and smooth random probabilities: many small regions, borders touched, several thresholds active"""
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    lab = (ndimage.gaussian_filter(rng.normal(size=(n, h, w)), (0, 2.0, 2.0)) > 0.22).astype(np.float32)
    prd = ndimage.gaussian_filter(rng.normal(size=(n, h, w)), (0, 1.5, 1.5))
    prd = (1 / (1 + np.exp(-(prd * 8 + 2.0 * lab - 1.0)))).astype(np.float32)
    lab[0] = 0                                                                  # a healthy slice
    if n > 1:
        prd[1] = 0                                                              # a slice without predictions
    return lab, prd[..., None]


@pytest.mark.gpu
@pytest.mark.parametrize('shape,thr,resize,morph', [((5, 64, 64), [0.5], 1.0, 5), ((4, 61, 45), [0.2, 0.5, 0.8, 0.95], 1.0, 5),
                                                    ((3, 96, 80), [0.3, 0.6], 0.5, 5), ((3, 50, 50), [0.5, 0.4], 1.0, 3),
                                                    ((2, 40, 72), [0.5], 1.0, 1), ((2, 48, 48), [0.5], 0.75, 4)])
def test_cuda_region_counts_match_oracle(shape, thr, resize, morph):
    from dnncancerannotator_b200.utils import metrics as M
    lab, prd = _random_case(sum(shape), *shape)
    want = rr.get_tp_fn_fp(lab, prd, thr, 0.3, resize, morph, raw=True)        # [B, 4, T]
    m = M.RegionBasedConfusionMatrix(thresholds=thr, IoU_threshold=0.3, resize_factor=resize, morph_filter_size=morph)
    tp, fn, fp = m.get_tp_fn_fp(lab, prd, None, return_raw=True)
    assert np.array_equal(tp, want[:, 0]) and np.array_equal(fn, want[:, 1]) and np.array_equal(fp, want[:, 2])
    tpp, fp2 = m.get_tp_fp(torch.from_numpy(lab).cuda(), torch.from_numpy(prd).cuda(), None)
    assert np.array_equal(tpp, want[:, 3].sum(0)) and np.array_equal(fp2, want[:, 2].sum(0))
    assert want[:, :3].sum() > 0                                                # the case is not vacuous
    # update_state accumulates; reset_state clears
    m.update_state(lab, prd)
    m.update_state(lab, prd)
    d = m.result_dict()
    assert L(d['true_positive_counts']) == L(2 * want[:, 0].sum(0)) and L(d['false_positive_counts']) == L(2 * want[:, 2].sum(0))
    m.reset_state()
    assert L(m.result_dict()['false_negative_counts']) == [0] * len(thr)


@pytest.mark.gpu
def test_cuda_ops_match_oracle_ops():
    """dnnca_connected_components / dnnca_grey_open / dnnca_resize_bilinear one by one"""
    from dnncancerannotator_b200 import native as N
    masks = _blobs(5, 4, 67, 93, 0.52)
    md = torch.from_numpy(masks.astype(np.uint8)).cuda()
    roots = torch.empty(masks.shape, dtype=torch.int32, device='cuda')
    N.call('dnnca_connected_components', N.stream_ptr(), N.ptr(md), *masks.shape, N.ptr(roots))
    roots = roots.cpu().numpy()
    ids = rr.connected_components(masks)
    base = 0
    for b in range(len(masks)):
        r = roots[b]
        assert np.array_equal(r >= 0, masks[b])
        uniq = np.unique(r[r >= 0])                          # ascending roots = raster order of first pixels
        rank = np.searchsorted(uniq, r) + 1 + base
        assert np.array_equal(np.where(r >= 0, rank, 0), ids[b])
        ys, xs = np.nonzero(r >= 0)
        assert np.all(r[ys, xs] <= ys * masks.shape[2] + xs)  # a root is the first pixel of its component
        base += len(uniq)
    rng = np.random.default_rng(6)
    p = rng.uniform(size=(3, 45, 70)).astype(np.float32)
    pd = torch.from_numpy(p).cuda()
    for k in (1, 2, 3, 5, 8, 15):
        out = torch.empty_like(pd)
        N.call('dnnca_grey_open', N.stream_ptr(), N.ptr(pd), *p.shape, k, N.ptr(out))
        assert np.array_equal(out.cpu().numpy(), rr.morph_open(p, k)), k
        for t in (0.3, 0.7):                                  # thresholding commutes with the opening
            assert np.array_equal(out.cpu().numpy() >= np.float32(t), rr.morph_open((p >= np.float32(t)).astype(np.int8), k) > 0)
    for oh, ow in [(22, 35), (45, 70), (13, 17), (60, 99)]:
        out = torch.empty(3, oh, ow, dtype=torch.float32, device='cuda')
        N.call('dnnca_resize_bilinear', N.stream_ptr(), N.ptr(pd), *p.shape, N.ptr(out), oh, ow)
        assert np.array_equal(out.cpu().numpy(), rr.resize_bilinear(p[..., None], oh, ow)[..., 0]), (oh, ow)


@pytest.mark.gpu
def test_cuda_pair_table_overflow_is_loud():
    """a checkerboard of isolated label pixels under one big prediction region: one pair per label pixel"""
    from dnncancerannotator_b200 import native as N
    from dnncancerannotator_b200.utils import metrics as M
    lab = np.zeros((1, 64, 64), np.float32)
    lab[0, ::2, ::2] = 1                                     # 1024 single-pixel regions
    prd = np.ones((1, 64, 64, 1), np.float32)
    m = M.RegionBasedConfusionMatrix(thresholds=0.5, table_slots=64)
    with pytest.raises(N.DnncaError, match='table_slots'):
        m.get_tp_fn_fp(lab, prd, None)
    m = M.RegionBasedConfusionMatrix(thresholds=0.5, table_slots=2048)
    tp, fn, fp = m.get_tp_fn_fp(lab, prd, None)
    assert (int(tp[0]), int(fn[0]), int(fp[0])) == (0, 1024, 1)


@pytest.mark.gpu
def test_cuda_full_size_properties():
    """256x256, batch 32, 10 thresholds (no oracle at this size): a prediction equal to the label detects every label
    region at every threshold <= 1 and raises no false positive; counts are invariant under a left-right flip."""
    from dnncancerannotator_b200.synthetic import make_slices
    from dnncancerannotator_b200.utils import metrics as M
    from scipy import ndimage
    _, y = make_slices(32, 256, 256, 3, seed=5)
    thr = thresholds_of(10)
    m = M.RegionBasedConfusionMatrix(thresholds=thr, IoU_threshold=0.3)
    tp, fn, fp = m.get_tp_fn_fp(y, y[..., None].astype(np.float32), None)
    n_regions = sum(ndimage.label(s > 0.5)[1] for s in y)
    assert n_regions > 0 and L(tp) == [n_regions] * 10 and L(fn) == [0] * 10 and L(fp) == [0] * 10
    rng = np.random.default_rng(0)
    p = ndimage.gaussian_filter(rng.normal(size=y.shape), (0, 3, 3))
    p = (1 / (1 + np.exp(-(p * 12 + 3 * y - 1.5)))).astype(np.float32)
    a = m.get_tp_fn_fp(y, p[..., None], None)
    b = m.get_tp_fn_fp(y[:, :, ::-1].copy(), p[:, :, ::-1, None].copy(), None)
    assert all(L(u) == L(v) for u, v in zip(a, b)) and sum(L(a[0])) + sum(L(a[2])) > 0
