"""CPU oracle for the per-frame skin-weight lookup -- TEST INFRASTRUCTURE ONLY.

Restates /root/reference/src/utils/gaussian_utils.py:167-196 ``skinning_weights_from_voxel_grid`` (called every step by
``HandGaussianModel.get_skin_weights``, src/models/hand_gaussian.py:65-76):
  xyz_norm = (xyz - grid_center) / grid_scale                                   (:171)
  grid_sample(grid_weights[D,H,W,C] as [1,C,D,H,W], xyz_norm, bilinear, zeros, align_corners=True)   (:173-179)
      -> xyz_norm[:,0] indexes W, [:,1] indexes H, [:,2] indexes D; index = (coord + 1) / 2 * (size - 1);
         8-corner trilinear interpolation, corners outside the grid contribute 0
  skin_wts = w / w.sum(-1)                                                      (:183)
plus the gradients autograd gives w.r.t. xyz and grid_weights.  Pinned: tests/test_oracle_skin.py checks this file against
tests/golden/skin_golden.npz (produced by the reference function itself, tests/golden/make_golden_skin.py).  numpy only.
"""
from __future__ import annotations

import numpy as np


def _corners(xyz, center, scale, shape, dtype):
    D, H, W, _ = shape
    n = (xyz.astype(dtype) - center.astype(dtype).reshape(1, 3)) / scale.astype(dtype).reshape(1, 3)
    size = np.array([W, H, D], dtype=dtype)
    pos = (n + 1) / 2 * (size - 1)                       # [N,3] as (ix, iy, iz)
    base = np.floor(pos)
    frac = pos - base
    return base.astype(np.int64), frac, size


def skin_weights(xyz, grid_center, grid_scale, grid_weights, dtype=np.float64):
    """-> (skin_wts [N,C], raw interpolated weights [N,C])."""
    D, H, W, C = grid_weights.shape
    base, frac, _ = _corners(xyz, grid_center, grid_scale, grid_weights.shape, dtype)
    raw = np.zeros((xyz.shape[0], C), dtype)
    g = grid_weights.astype(dtype)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                ix, iy, iz = base[:, 0] + dx, base[:, 1] + dy, base[:, 2] + dz
                wgt = (frac[:, 0] if dx else 1 - frac[:, 0]) * (frac[:, 1] if dy else 1 - frac[:, 1]) * (frac[:, 2] if dz else 1 - frac[:, 2])
                ok = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H) & (iz >= 0) & (iz < D)
                vals = np.zeros_like(raw)
                vals[ok] = g[iz[ok], iy[ok], ix[ok]]
                raw += wgt[:, None] * vals
    with np.errstate(invalid="ignore", divide="ignore"):
        return raw / raw.sum(-1, keepdims=True), raw


def skin_weights_backward(xyz, grid_center, grid_scale, grid_weights, g_out, dtype=np.float64):
    """Gradients of sum(skin_wts * g_out) -> (g_xyz [N,3], g_grid [D,H,W,C])."""
    D, H, W, C = grid_weights.shape
    base, frac, size = _corners(xyz, grid_center, grid_scale, grid_weights.shape, dtype)
    wts, raw = skin_weights(xyz, grid_center, grid_scale, grid_weights, dtype)
    S = raw.sum(-1, keepdims=True)
    go = g_out.astype(dtype)
    g_raw = (go - (go * wts).sum(-1, keepdims=True)) / S           # through the row normalisation
    g = grid_weights.astype(dtype)
    g_pos = np.zeros((xyz.shape[0], 3), dtype)
    g_grid = np.zeros(grid_weights.shape, dtype)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                ix, iy, iz = base[:, 0] + dx, base[:, 1] + dy, base[:, 2] + dz
                fx = frac[:, 0] if dx else 1 - frac[:, 0]
                fy = frac[:, 1] if dy else 1 - frac[:, 1]
                fz = frac[:, 2] if dz else 1 - frac[:, 2]
                ok = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H) & (iz >= 0) & (iz < D)
                vals = np.zeros_like(raw)
                vals[ok] = g[iz[ok], iy[ok], ix[ok]]
                dot = (vals * g_raw).sum(-1)
                g_pos[:, 0] += (1 if dx else -1) * fy * fz * dot
                g_pos[:, 1] += (1 if dy else -1) * fx * fz * dot
                g_pos[:, 2] += (1 if dz else -1) * fx * fy * dot
                np.add.at(g_grid, (iz[ok], iy[ok], ix[ok]), (fx * fy * fz)[ok, None] * g_raw[ok])
    g_xyz = g_pos * ((size - 1) / 2) / grid_scale.astype(dtype).reshape(1, 3)
    return g_xyz, g_grid
