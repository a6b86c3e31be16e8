"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the thin-plate-spline warp augmentation,
``random_warp`` (annotator/data.py:718-763) -> ``tfa.image.sparse_image_warp`` (SURVEY 8f "later" row).

The arithmetic lives in tensorflow-addons (unpinned, requirements.txt:3; not vendored, not installable here), in three
modules whose published algorithm is restated below: ``image/sparse_image_warp.py`` (grid of query locations, call
order), ``image/interpolate_spline.py`` (polyharmonic spline: solve [[A, B], [B^T, 0]] [w; v] = [f; 0] with
A_ij = phi(|c_i - c_j|^2), B = [c, 1]; evaluate phi(|q - c_i|^2) w + [q, 1] v; phi(r) = 0.5 r log(max(r, 1e-10)) for
order 2, r being the SQUARED distance) and ``image/dense_image_warp.py`` (bilinear resampling at grid - flow with
floor clamped to [0, size-2] and weights to [0, 1]).  The reference has no test or fixture for this path and TensorFlow
cannot run here: parity of this module is pinned only against an independent implementation of the same interpolant,
``scipy.interpolate.RBFInterpolator(kernel='thin_plate_spline', degree=1)`` (r^2 log r = 0.5 r^2 log r^2, linear
polynomial tail), and against definition-level known answers (tests/test_tps_warp.py): "parity unpinned" in the sense
of the task statement.

``dtype`` selects the working precision: float64 = the mathematical answer the CUDA path is held to; float32 = the
precision the reference runs in (tf.float32 throughout, distances by the |x|^2 - 2xy + |y|^2 expansion), used to
measure how far the reference's own rounding sits from that answer.
"""
import numpy as np

EPSILON = 0.0000000001


def _phi(r, order):
    """interpolate_spline._phi; ``r`` is a squared distance."""
    dt = r.dtype.type
    if order == 1:
        return np.sqrt(np.maximum(r, dt(EPSILON)))
    if order == 2:
        return dt(0.5) * r * np.log(np.maximum(r, dt(EPSILON)))
    if order == 4:
        return dt(0.5) * np.square(r) * np.log(np.maximum(r, dt(EPSILON)))
    if order % 2 == 0:
        r = np.maximum(r, dt(EPSILON))
        return dt(0.5) * np.power(r, dt(0.5 * order)) * np.log(r)
    r = np.maximum(r, dt(EPSILON))
    return np.power(r, dt(0.5 * order))


def _cross_squared_distance_matrix(x, y):
    """interpolate_spline._cross_squared_distance_matrix: [b,n,d],[b,m,d] -> [b,n,m] by the norm expansion."""
    xn = np.sum(np.square(x), 2)
    yn = np.sum(np.square(y), 2)
    return xn[:, :, None] - 2 * np.matmul(x, np.swapaxes(y, 1, 2)) + yn[:, None, :]


def _pairwise_squared_distance_matrix(x):
    xx = np.matmul(x, np.swapaxes(x, 1, 2))
    xn = np.diagonal(xx, axis1=1, axis2=2)
    return xn[:, :, None] - 2 * xx + xn[:, None, :]


def solve_interpolation(train_points, train_values, order=2, regularization_weight=0.0):
    """interpolate_spline._solve_interpolation -> (w [b,n,k], v [b,d+1,k])"""
    c, f = train_points, train_values
    b, n, d = c.shape
    k = f.shape[-1]
    a = _phi(_pairwise_squared_distance_matrix(c), order)
    if regularization_weight > 0:
        a = a + c.dtype.type(regularization_weight) * np.eye(n, dtype=c.dtype)[None]
    bm = np.concatenate([c, np.ones_like(c[..., :1])], 2)                       # [b,n,d+1]
    left = np.concatenate([a, np.swapaxes(bm, 1, 2)], 1)                        # [b,n+d+1,n]
    right = np.concatenate([bm, np.zeros((b, d + 1, d + 1), c.dtype)], 1)       # [b,n+d+1,d+1]
    lhs = np.concatenate([left, right], 2)
    rhs = np.concatenate([f, np.zeros((b, d + 1, k), c.dtype)], 1)
    wv = np.linalg.solve(lhs, rhs).astype(c.dtype)
    return wv[:, :n], wv[:, n:]


def apply_interpolation(query_points, train_points, w, v, order=2):
    """interpolate_spline._apply_interpolation"""
    ph = _phi(_cross_squared_distance_matrix(query_points, train_points), order)
    q1 = np.concatenate([query_points, np.ones_like(query_points[..., :1])], 2)
    return np.matmul(ph, w) + np.matmul(q1, v)


def interpolate_spline(train_points, train_values, query_points, order=2, regularization_weight=0.0):
    w, v = solve_interpolation(train_points, train_values, order, regularization_weight)
    return apply_interpolation(query_points, train_points, w, v, order)


def interpolate_bilinear(grid, query_points):
    """dense_image_warp.interpolate_bilinear, indexing='ij': grid [b,h,w,c], query_points [b,n,2] -> [b,n,c]"""
    b, h, w, c = grid.shape
    gt = grid.dtype.type
    floors, ceils, alphas = [], [], []
    for dim, size in ((0, h), (1, w)):
        q = query_points[..., dim]
        fl = np.minimum(np.maximum(q.dtype.type(0), np.floor(q)), q.dtype.type(size - 2))
        ifl = fl.astype(np.int32)
        floors.append(ifl)
        ceils.append(ifl + 1)
        al = (q - fl).astype(grid.dtype)
        alphas.append(np.minimum(np.maximum(gt(0), al), gt(1))[..., None])
    bi = np.arange(b)[:, None]
    tl, tr = grid[bi, floors[0], floors[1]], grid[bi, floors[0], ceils[1]]
    bl, br = grid[bi, ceils[0], floors[1]], grid[bi, ceils[0], ceils[1]]
    top = alphas[1] * (tr - tl) + tl
    bot = alphas[1] * (br - bl) + bl
    return alphas[0] * (bot - top) + top


def dense_image_warp(image, flow):
    """dense_image_warp.dense_image_warp: out[b,y,x] = image[b, y - flow[b,y,x,0], x - flow[b,y,x,1]] (bilinear)"""
    b, h, w, c = image.shape
    gy, gx = np.meshgrid(np.arange(h), np.arange(w), indexing='ij')
    grid = np.stack([gy, gx], 2).astype(flow.dtype)[None]
    q = (grid - flow).reshape(b, h * w, 2)
    return interpolate_bilinear(image, q).reshape(b, h, w, c)


def sparse_image_warp(image, source_control_point_locations, dest_control_point_locations, interpolation_order=2,
                      regularization_weight=0.0, dtype=np.float64):
    """sparse_image_warp.sparse_image_warp with num_boundary_points=0 (the reference's call, data.py:749-753)
    -> (warped image [b,h,w,c], dense flow [b,h,w,2]); control points are (row, column)."""
    image = np.asarray(image, dtype)
    src = np.asarray(source_control_point_locations, dtype)
    dst = np.asarray(dest_control_point_locations, dtype)
    b, h, w, _ = image.shape
    flows = dst - src
    gy, gx = np.meshgrid(np.linspace(0, h - 1, h), np.linspace(0, w - 1, w), indexing='ij')
    grid = np.broadcast_to(np.stack([gy, gx], -1).reshape(1, h * w, 2).astype(dtype), (b, h * w, 2))
    flat = interpolate_spline(dst, flows, grid, interpolation_order, regularization_weight)
    dense = flat.reshape(b, h, w, 2)
    return dense_image_warp(image, dense), dense


def draw_control_points(rng, n_images, width, n_points=100, max_diff=5, stddev=2.0):
    """data.py:742-746: raw ~ U[0, width)^2, diff ~ clip(N(0, stddev), -max_diff, max_diff); (source, dest) float32"""
    raw = rng.uniform(0.0, float(width), (n_images, n_points, 2)).astype(np.float32)
    diff = np.clip(rng.normal(0.0, stddev, (n_images, n_points, 2)), -float(max_diff), float(max_diff)).astype(np.float32)
    return raw, raw + diff
