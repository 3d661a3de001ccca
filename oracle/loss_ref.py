"""CPU oracle for the image losses of the training step -- TEST INFRASTRUCTURE ONLY (imported by tests/ and bench.py's CPU
baseline leg; the product never imports it).

Restates /root/reference/src/utils/loss_utils.py:22-97 as the reference CALLS it (src/modules/base.py:323-365):

  * ``l1_loss(pred, gt, mean=False).mean()``                                   loss_utils.py:22-27, base.py:329-331
  * ``ssim(pred, gt)`` with pred [H,W,3] and gt [1,H,W,3]                      loss_utils.py:57-97, base.py:347
    ``channel = img1.size(-3)`` (:58) is H for these HWC tensors, so ``F.conv2d(..., groups=channel)`` (:69-83) filters
    every image ROW independently: the 11x11 Gaussian window (sigma 1.5, :38-54) slides over the (W, 3) plane of the row
    with zero padding 5.  Along W that is an 11-tap filter; along the 3-wide colour axis the window reaches all three
    channels, i.e. a 3x3 mixing matrix M[c][c'] = g[5 + c' - c].  There is no vertical filtering.
  * final loss = 0.8 * L1 + 0.2 * (1 - SSIM)                                   config/{OBJ_GAUSSIAN,COMPOSITE}.yaml:22-23

Pinned: tests/test_oracle_loss.py checks these functions against tests/golden/loss_golden.npz, which was produced by the
reference's own loss_utils (tests/golden/make_golden_loss.py).  numpy only; float64 accumulation optional.
"""
from __future__ import annotations

import math

import numpy as np

C1, C2 = 0.01 ** 2, 0.03 ** 2          # loss_utils.py:85-86
WINDOW, SIGMA = 11, 1.5                # loss_utils.py:57, 48


def gaussian_taps(dtype=np.float32) -> np.ndarray:
    """loss_utils.py:38-45: normalised 11-tap Gaussian (computed in Python floats, stored as fp32 like torch.Tensor)."""
    g = np.array([math.exp(-((x - WINDOW // 2) ** 2) / float(2 * SIGMA ** 2)) for x in range(WINDOW)], dtype=np.float32)
    total = np.float32(g.astype(np.float64).sum())      # torch's fp32 sum of these 11 values rounds like the exact sum
    return (g / total).astype(dtype)


def window_factors(dtype=np.float32):
    """(taps along W [11], mixing matrix over the colour axis [3,3]) of the 2-D window g g^T (loss_utils.py:48-54).
    window[i][j] = g[i] * g[j] in fp32, so filtering = M applied across channels, then the taps along W."""
    g = gaussian_taps(dtype)
    M = np.array([[g[5 + cp - c] for cp in range(3)] for c in range(3)], dtype=dtype)
    return g, M


def _filter(img: np.ndarray, dtype) -> np.ndarray:
    """conv2d(img, window, padding=5, groups=H) of an [H,W,3] tensor (loss_utils.py:69): per row, zero padded."""
    g, M = window_factors(dtype)
    H, W, _ = img.shape
    mixed = img.astype(dtype) @ M.T                      # out[.., c] = sum_c' M[c][c'] img[.., c']
    pad = np.zeros((H, W + 10, 3), dtype)
    pad[:, 5:5 + W] = mixed
    out = np.zeros((H, W, 3), dtype)
    for k in range(WINDOW):
        out += g[k] * pad[:, k:k + W]
    return out


def ssim_map(pred: np.ndarray, gt: np.ndarray, dtype=np.float32):
    """loss_utils.py:68-91 -> (ssim_map [H,W,3], intermediates for the gradient)."""
    p, q = pred.astype(dtype), gt.astype(dtype)
    mu1, mu2 = _filter(p, dtype), _filter(q, dtype)
    e11, e22, e12 = _filter(p * p, dtype), _filter(q * q, dtype), _filter(p * q, dtype)
    s1, s2, s12 = e11 - mu1 * mu1, e22 - mu2 * mu2, e12 - mu1 * mu2
    A1, A2 = 2 * mu1 * mu2 + dtype(C1), 2 * s12 + dtype(C2)
    B1, B2 = mu1 * mu1 + mu2 * mu2 + dtype(C1), s1 + s2 + dtype(C2)
    return (A1 * A2) / (B1 * B2), dict(mu1=mu1, mu2=mu2, A1=A1, A2=A2, B1=B1, B2=B2)


def photometric_loss(pred: np.ndarray, gt: np.ndarray, w_l1: float = 0.8, w_ssim: float = 0.2, dtype=np.float32):
    """-> dict(l1, ssim, loss, grad [H,W,3]) of  w_l1 * mean|pred - gt| + w_ssim * (1 - mean(ssim_map))  (base.py:323-365)."""
    p, q = pred.astype(dtype), gt.astype(dtype)
    n = p.size
    smap, t = ssim_map(p, q, dtype)
    l1 = np.abs(p - q).mean(dtype=np.float64)
    ss = smap.mean(dtype=np.float64)
    # d ssim / d (mu1, E11, E12); sigma1^2 = E11 - mu1^2, sigma12 = E12 - mu1 mu2
    mu1, mu2, A1, A2, B1, B2 = t["mu1"], t["mu2"], t["A1"], t["A2"], t["B1"], t["B2"]
    den = B1 * B2
    d_mu1 = ((2 * mu2 * A2 - 2 * mu2 * A1) * den - A1 * A2 * (2 * mu1 * B2 - 2 * mu1 * B1)) / (den * den)
    d_e11 = -(A1 * A2) / (B1 * B2 * B2)
    d_e12 = 2 * A1 / den
    up = dtype(-w_ssim / n)                                 # d loss / d ssim_map
    # the filter is self-adjoint (symmetric taps, symmetric mixing matrix, zero padding)
    g_pred = _filter(up * d_mu1, dtype) + 2 * p * _filter(up * d_e11, dtype) + q * _filter(up * d_e12, dtype)
    g_pred = g_pred + dtype(w_l1 / n) * np.sign(p - q)
    return dict(l1=float(l1), ssim=float(ss), loss=float(w_l1 * l1 + w_ssim * (1.0 - ss)), grad=g_pred)
