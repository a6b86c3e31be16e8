"""Thin-plate-spline warp augmentation (SURVEY 8f "later" row; annotator/data.py:718-763 -> tfa.image.sparse_image_warp).

CPU: the numpy oracle (``oracle/ref_warp.py``) against an independent implementation of the same interpolant
(scipy's RBFInterpolator) and against definition-level known answers; how far the reference's OWN float32 arithmetic
sits from the exact interpolant (reported, bounds the meaning of any tolerance).  GPU: ``dnnca_tps_fit`` +
``dnnca_tps_warp`` against the float64 oracle.

Tolerances (floating point; BASELINE's north star states none for this row): dense flow within 1e-3 px of the float64
oracle (the reference's float32 graph is 0.1-0.3 px away from it), warped image within 1e-3 * value range + flow
tolerance * local gradient, stated per assertion.
"""
import numpy as np
import pytest
import torch

from oracle import ref_warp as rw


def _smooth_image(rng, n, size, c):
    from scipy import ndimage
    img = ndimage.gaussian_filter(rng.uniform(size=(n, size, size, c)), (0, 2.0, 2.0, 0))
    img = (img - img.min()) / (img.max() - img.min())
    return img.astype(np.float32)


def test_oracle_matches_scipy_thin_plate_spline():
    from scipy.interpolate import RBFInterpolator
    rng = np.random.default_rng(0)
    n, size, npts = 2, 48, 60
    src, dst = rw.draw_control_points(rng, n, size, npts, 5, 2.0)
    img = _smooth_image(rng, n, size, 2)
    _, flow = rw.sparse_image_warp(img, src, dst, dtype=np.float64)
    gy, gx = np.meshgrid(np.arange(size), np.arange(size), indexing='ij')
    q = np.stack([gy, gx], -1).reshape(-1, 2).astype(np.float64)
    for b in range(n):
        ref = RBFInterpolator(dst[b].astype(np.float64), (dst[b] - src[b]).astype(np.float64), kernel='thin_plate_spline',
                              degree=1)(q).reshape(size, size, 2)
        assert np.abs(ref - flow[b]).max() < 1e-6


def test_oracle_known_answers():
    rng = np.random.default_rng(1)
    size = 32
    img = _smooth_image(rng, 1, size, 3).astype(np.float64)
    src, _ = rw.draw_control_points(rng, 1, size, 20)
    # no displacement: identity
    out, flow = rw.sparse_image_warp(img, src, src)
    assert np.abs(flow).max() < 1e-9 and np.abs(out - img).max() < 1e-9
    # every control point moved by (2, -3): the affine part carries it, flow is constant, image shifts by whole pixels
    out, flow = rw.sparse_image_warp(img, src, src + np.array([2.0, -3.0], np.float32))
    assert np.abs(flow - np.array([2.0, -3.0])).max() < 1e-5                        # float32 control points
    assert np.abs(out[0, 4:-4, 4:-4] - img[0, 2:-6, 7:-1]).max() < 1e-5            # out[y, x] = img[y - 2, x + 3]
    # an affine displacement field is reproduced exactly (polyharmonic splines of order 2 contain the linear polynomials)
    A = np.array([[0.02, -0.01], [0.015, 0.03]])
    dst = (src.astype(np.float64) @ (np.eye(2) + A).T + np.array([0.5, -0.25])).astype(np.float32)
    _, flow = rw.sparse_image_warp(img, src, dst)
    gy, gx = np.meshgrid(np.arange(size), np.arange(size), indexing='ij')
    q = np.stack([gy, gx], -1).astype(np.float64)
    # flow is a function of the DEST location: f(d) = d - s with s = (I+A)^-1 (d - t)
    Minv = np.linalg.inv(np.eye(2) + A)
    want = q - (q - np.array([0.5, -0.25])) @ Minv.T
    assert np.abs(flow[0] - want).max() < 2e-5                                       # float32 control points
    # interpolation property: control points on grid nodes get exactly their displacement
    nodes = np.stack(np.meshgrid(np.arange(4, 32, 8), np.arange(4, 32, 8), indexing='ij'), -1).reshape(1, -1, 2).astype(np.float32)
    disp = rng.normal(0, 1.5, nodes.shape).astype(np.float32)
    _, flow = rw.sparse_image_warp(img, nodes - disp, nodes)
    got = flow[0][nodes[0, :, 0].astype(int), nodes[0, :, 1].astype(int)]
    assert np.abs(got - disp[0]).max() < 1e-6
    # bilinear clamping: a flow that points outside samples the border rows (floor <= size-2, weight <= 1)
    far = rw.dense_image_warp(img, np.full((1, size, size, 2), -100.0))
    assert np.abs(far - img[:, -1:, -1:, :]).max() < 1e-12


def test_reference_precision_noise_is_reported():
    """the reference's float32 graph (norm-expansion distances) against the exact interpolant: the yardstick"""
    rng = np.random.default_rng(2)
    src, dst = rw.draw_control_points(rng, 1, 128, 100, 5, 2.0)
    img = _smooth_image(rng, 1, 128, 1)
    _, f64 = rw.sparse_image_warp(img, src, dst, dtype=np.float64)
    _, f32 = rw.sparse_image_warp(img, src, dst, dtype=np.float32)
    noise = float(np.abs(f32 - f64).max())
    assert 1e-3 < noise < 5.0, noise         # order 0.1 px: any tolerance tighter than this is tighter than the reference itself


@pytest.mark.gpu
@pytest.mark.parametrize('n,size,c,npts,max_diff,stddev', [(3, 64, 3, 50, 5, 3.0), (2, 256, 6, 100, 5, 2.0),
                                                            (2, 200, 1, 150, 15, 20.0), (1, 96, 5, 7, 5, 2.0),
                                                            (2, 128, 2, 200, 5, 2.0)])      # 200 points: the system leaves shared memory
def test_cuda_warp_matches_float64_oracle(n, size, c, npts, max_diff, stddev):
    from dnncancerannotator_b200 import data_tail as DT
    rng = np.random.default_rng(n * 1000 + size + npts)
    img = _smooth_image(rng, n, size, c)
    src, dst = rw.draw_control_points(rng, n, size, npts, max_diff, stddev)
    want, wflow = rw.sparse_image_warp(img, src, dst, dtype=np.float64)
    got, gflow = DT.sparse_image_warp(img, src, dst)
    got, gflow = got.cpu().numpy(), gflow.cpu().numpy()
    ferr = float(np.abs(gflow - wflow).max())
    assert ferr < 1e-3, ferr                                                          # pixels
    # image: bilinear of a smooth image is Lipschitz in the sample position; gradient bound from the image itself
    grad = max(float(np.abs(np.diff(img, axis=1)).max()), float(np.abs(np.diff(img, axis=2)).max()))
    assert float(np.abs(got - want).max()) < 1e-5 + 2 * grad * 1e-3
    assert np.isfinite(got).all() and got.min() >= img.min() - 1e-6 and got.max() <= img.max() + 1e-6


@pytest.mark.gpu
def test_cuda_warp_known_answers_and_api():
    from dnncancerannotator_b200 import data_tail as DT
    from dnncancerannotator_b200 import native as N
    rng = np.random.default_rng(3)
    img = _smooth_image(rng, 2, 64, 4)
    src, _ = rw.draw_control_points(rng, 2, 64, 30)
    out, flow = DT.sparse_image_warp(img, src, src)                                   # identity
    assert float(flow.abs().max()) < 1e-5 and float((out.cpu() - torch.from_numpy(img)).abs().max()) < 1e-5
    out, flow = DT.sparse_image_warp(img, src, src + np.array([2.0, -3.0], np.float32))
    assert float((flow.cpu() - torch.tensor([2.0, -3.0])).abs().max()) < 1e-4
    assert float((out.cpu()[:, 4:-4, 4:-4] - torch.from_numpy(img)[:, 2:-6, 7:-1]).abs().max()) < 1e-4
    with pytest.raises(N.DnncaError, match='singular'):
        dup = src.copy()
        dup[0, 1] = dup[0, 0]
        DT.sparse_image_warp(img, dup, dup + 1.0)
    with pytest.raises(NotImplementedError):
        DT.sparse_image_warp(img, src, src, interpolation_order=3)
    # random_warp: single image and batch forms (data.py:733-738), label channel stays inside [0, 1]
    one = DT.random_warp(img[0], n_points=50, max_diff=5, stddev=3, process_in_batch=None, rng=np.random.default_rng(0))
    assert tuple(one.shape) == (64, 64, 4) and float(one.min()) >= 0.0 and float(one.max()) <= 1.0
    two = DT.random_warp(img, process_in_batch=2, rng=np.random.default_rng(0))
    assert tuple(two.shape) == (2, 64, 64, 4) and float((two.cpu() - torch.from_numpy(img)).abs().max()) > 1e-3
    with pytest.raises(ValueError):
        DT.random_warp(img[:, :32], process_in_batch=2)
