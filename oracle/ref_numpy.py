"""Second, loop-level numpy-fp64 restatement (forward AND hand-derived backward).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``); parity unpinned.

Written directly from the defining formulas (SURVEY.md 8a "Keras semantics"),
with explicit index loops and no library convolution, so that it can arbitrate
between ``ref_ops`` (torch/oneDNN + autograd) and the CUDA kernels on tiny
shapes.  The backward functions are the formulas the CUDA dgrad/wgrad/bwd
kernels implement.
"""
import numpy as np

F64 = np.float64


# ---- Conv2D 'same', stride 1 (components.py:47-50) ---------------------------
def conv2d_same_fwd(x, k, b=None):
    """y[n,i,j,co] = sum_{a,c,ci} x[n,i+a-ph,j+c-pw,ci] * k[a,c,ci,co] + b[co] (zeros outside)."""
    x, k = x.astype(F64), k.astype(F64)
    n, h, w, cin = x.shape
    kh, kw, _, cout = k.shape
    ph, pw = kh // 2, kw // 2
    y = np.zeros((n, h, w, cout), F64)
    for i in range(h):
        for j in range(w):
            for a in range(kh):
                for c in range(kw):
                    ii, jj = i + a - ph, j + c - pw
                    if 0 <= ii < h and 0 <= jj < w:
                        y[:, i, j, :] += x[:, ii, jj, :] @ k[a, c]
    if b is not None:
        y += b.astype(F64)
    return y


def conv2d_same_bwd(x, k, dy):
    """dx[n,p,q,ci] = sum dy[n,p-a+ph,q-c+pw,co]*k[a,c,ci,co];
    dk[a,c,ci,co] = sum x[n,i+a-ph,j+c-pw,ci]*dy[n,i,j,co]; db[co] = sum dy."""
    x, k, dy = x.astype(F64), k.astype(F64), dy.astype(F64)
    n, h, w, cin = x.shape
    kh, kw, _, cout = k.shape
    ph, pw = kh // 2, kw // 2
    dx = np.zeros_like(x)
    dk = np.zeros_like(k)
    for i in range(h):
        for j in range(w):
            for a in range(kh):
                for c in range(kw):
                    ii, jj = i + a - ph, j + c - pw
                    if 0 <= ii < h and 0 <= jj < w:
                        dx[:, ii, jj, :] += dy[:, i, j, :] @ k[a, c].T
                        dk[a, c] += x[:, ii, jj, :].T @ dy[:, i, j, :]
    return dx, dk, dy.sum(axis=(0, 1, 2))


# ---- Conv2DTranspose k = s = 2 (components.py:118-120) ------------------------
def tconv2x2_fwd(x, k, b):
    """out[n,2i+a,2j+c,co] = sum_ci x[n,i,j,ci]*k[a,c,co,ci] + b[co]."""
    x, k = x.astype(F64), k.astype(F64)
    n, h, w, cin = x.shape
    s = k.shape[0]
    cout = k.shape[2]
    y = np.zeros((n, s * h, s * w, cout), F64)
    for a in range(s):
        for c in range(s):
            y[:, a::s, c::s, :] = x @ k[a, c].T
    return y + b.astype(F64)


def tconv2x2_bwd(x, k, dy):
    x, k, dy = x.astype(F64), k.astype(F64), dy.astype(F64)
    s = k.shape[0]
    dx = np.zeros_like(x)
    dk = np.zeros_like(k)
    for a in range(s):
        for c in range(s):
            g = dy[:, a::s, c::s, :]                      # [n,h,w,co]
            dx += g @ k[a, c]                               # [co,ci]
            dk[a, c] = np.einsum('nhwo,nhwi->oi', g, x)
    return dx, dk, dy.sum(axis=(0, 1, 2))


# ---- MaxPool 2x2/2 (components.py:54) -----------------------------------------
def maxpool2x2_fwd(x):
    """First maximum in row-major window order wins (strict '>' scan)."""
    n, h, w, c = x.shape
    y = np.empty((n, h // 2, w // 2, c), x.dtype)
    idx = np.zeros((n, h // 2, w // 2, c), np.uint8)
    for i in range(h // 2):
        for j in range(w // 2):
            best = x[:, 2 * i, 2 * j, :].copy()
            bi = np.zeros(best.shape, np.uint8)
            for t in (1, 2, 3):
                v = x[:, 2 * i + t // 2, 2 * j + t % 2, :]
                m = v > best
                best[m] = v[m]
                bi[m] = t
            y[:, i, j, :] = best
            idx[:, i, j, :] = bi
    return y, idx


def maxpool2x2_bwd(dy, idx):
    n, hh, wh, c = dy.shape
    dx = np.zeros((n, 2 * hh, 2 * wh, c), dy.dtype)
    for t in range(4):
        dx[:, t // 2::2, t % 2::2, :] = np.where(idx == t, dy, 0)
    return dx


# ---- BatchNorm, training mode (components.py:57) -----------------------------
def bn_train_fwd(x, gamma, beta, eps=1e-3):
    x = x.astype(F64)
    mean = x.mean(axis=(0, 1, 2))
    var = ((x - mean) ** 2).mean(axis=(0, 1, 2))
    invstd = 1.0 / np.sqrt(var + eps)
    xhat = (x - mean) * invstd
    g = np.ones_like(mean) if gamma is None else gamma.astype(F64)
    return xhat * g + beta.astype(F64), mean, var, invstd


def bn_train_bwd(x, gamma, dy, mean, invstd):
    """dgamma = sum dy*xhat ; dbeta = sum dy ;
    dx = gamma*invstd*(dy - dbeta/M - xhat*dgamma/M)."""
    x, dy = x.astype(F64), dy.astype(F64)
    m = x.shape[0] * x.shape[1] * x.shape[2]
    xhat = (x - mean) * invstd
    dgamma = (dy * xhat).sum(axis=(0, 1, 2))
    dbeta = dy.sum(axis=(0, 1, 2))
    g = np.ones_like(mean) if gamma is None else gamma.astype(F64)
    dx = g * invstd * (dy - dbeta / m - xhat * dgamma / m)
    return dx, dgamma, dbeta


# ---- head + loss (unet.py:241-244 ; losses.py:17-37) ---------------------------
def weighted_bce_fwd_bwd(label, logits, weight=None, weight_add=0.0, weight_mul=1.0, n_replicas=1):
    """Returns (per_sample[B], dlogits[B,H,W]) with dlogits = d(mean_B per_sample / R)/dz
    = mask*(sigmoid(z)-y)/(B*H*W*R)."""
    y, z = label.astype(F64), logits.astype(F64).reshape(label.shape)
    if weight is None:
        r = y.sum() / y.size
        weight = 1.0 / r if r > 0 else 1.0
    wgt = weight_mul * weight + weight_add
    mask = y * (wgt - 1.0) + 1.0
    bce = np.maximum(z, 0) - z * y + np.log1p(np.exp(-np.abs(z)))
    per_sample = (bce * mask).mean(axis=(1, 2))
    sig = 1.0 / (1.0 + np.exp(-z))
    dz = mask * (sig - y) / (y.size * n_replicas)
    return per_sample, dz


def act_bwd(y, g, act):
    """Gradient through relu / leaky-relu given the *output* y (sign-preserving)."""
    if act is None:
        return g
    if act == 'relu':
        return np.where(y > 0, g, 0.0)
    if isinstance(act, tuple) and act[0] == 'leaky':
        return np.where(y > 0, g, act[1] * g)
    raise ValueError(act)
