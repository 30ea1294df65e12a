"""Pure-numpy direct-loop restatement of the layer conventions.  TEST INFRASTRUCTURE ONLY.

Exists to guard the torch oracle (oracle/resuneta_oracle.py) against a shared
misunderstanding of padding / dilation / stride / pooling / BN conventions.
Small shapes only (pure Python loops).  NHWC, fp64.

Conventions restated (keras defaults; call sites model2.py:17-24,36-39,47-60):
  * Conv2D 'same', k=3, stride 1, dilation d: symmetric zero pad d, applied AFTER the
    BN+ReLU that precede it (model2.py:17-20).
  * Conv2D 1x1 stride 2 'valid': samples input pixels [0::2, 0::2].
  * MaxPooling2D(k): stride k, valid (floor).  UpSampling2D(k): nearest.
  * BatchNormalization training: biased batch variance, eps 1e-3.
"""
import numpy as np


def conv2d(x, w, b, stride=1, dilation=1, same=False):
    n, h, wd, cin = x.shape
    kh, kw, _, cout = w.shape
    pad = dilation * (kh - 1) // 2 if same else 0
    ho = (h + 2 * pad - dilation * (kh - 1) - 1) // stride + 1
    wo = (wd + 2 * pad - dilation * (kw - 1) - 1) // stride + 1
    y = np.zeros((n, ho, wo, cout), dtype=np.float64)
    for i in range(ho):
        for j in range(wo):
            acc = np.tile(b.astype(np.float64), (n, 1))
            for a in range(kh):
                for c in range(kw):
                    ii = i * stride - pad + a * dilation
                    jj = j * stride - pad + c * dilation
                    if 0 <= ii < h and 0 <= jj < wd:
                        acc += x[:, ii, jj, :].astype(np.float64) @ w[a, c].astype(np.float64)
            y[:, i, j, :] = acc
    return y


def bn_train(x, gamma, beta, eps=1e-3):
    c = x.shape[-1]
    flat = x.reshape(-1, c).astype(np.float64)
    mean = flat.sum(0) / flat.shape[0]
    var = ((flat - mean) ** 2).sum(0) / flat.shape[0]
    return (x - mean) / np.sqrt(var + eps) * gamma + beta, mean, var


def maxpool(x, k):
    n, h, w, c = x.shape
    ho, wo = h // k, w // k
    y = np.zeros((n, ho, wo, c), dtype=x.dtype)
    for i in range(ho):
        for j in range(wo):
            y[:, i, j, :] = x[:, i * k:(i + 1) * k, j * k:(j + 1) * k, :].max(axis=(1, 2))
    return y


def upsample(x, k):
    return np.repeat(np.repeat(x, k, axis=1), k, axis=2)


def tanimoto_dual(label, pred):
    """multitasking_utils.py:38-85 in explicit loops (fp64)."""
    def T(a, b):   # a = "label" slot (weights), b = "pred" slot
        bsz, h, w, c = a.shape
        v = np.zeros(c)
        for ci in range(c):
            v[ci] = a[..., ci].sum() / bsz
        with np.errstate(divide="ignore"):
            wl = 1.0 / v ** 2
        fin = np.where(np.isinf(wl), 0.0, wl)
        wl = np.where(np.isinf(wl), fin.max(), wl)
        out = np.zeros(bsz)
        for bi in range(bsz):
            num = den = 0.0
            for ci in range(c):
                sp = (a[bi, ..., ci] * b[bi, ..., ci]).sum()
                sq = (a[bi, ..., ci] ** 2).sum() + (b[bi, ..., ci] ** 2).sum()
                num += wl[ci] * sp
                den += wl[ci] * (sq - sp)
            out[bi] = (num + 1e-5) / (den + 1e-5)
        return out
    return 1.0 - 0.5 * (T(pred, label) + T(1.0 - label, 1.0 - pred))
