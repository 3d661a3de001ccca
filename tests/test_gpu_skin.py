"""Skin-weight lookup kernels (manus_b200/csrc/skin.cu) against the golden vectors of the reference's own
skinning_weights_from_voxel_grid and against the pinned oracle at a realistic grid size."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN
from oracle import skin_ref

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(GOLDEN, "skin_golden.npz"))


def run(xyz, center, scale, grid, gout, grid_grad=True):
    from manus_b200.skinning import skinning_weights_from_voxel_grid

    x = torch.tensor(xyz, device="cuda").requires_grad_(True)
    g = torch.tensor(grid, device="cuda").requires_grad_(grid_grad)
    w = skinning_weights_from_voxel_grid(x, torch.tensor(center), torch.tensor(scale), g)
    (w * torch.tensor(gout, device="cuda")).sum().backward()
    return w.detach().cpu().numpy(), x.grad.cpu().numpy(), None if not grid_grad else g.grad.cpu().numpy()


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_skin_weights_match_reference_golden(built_lib, name):
    a = {k[len(name) + 1:]: G[k] for k in G.files if k.startswith(name + "_")}
    w, g_xyz, g_grid = run(a["xyz"], a["center"], a["scale"], a["grid"], a["gout"])
    np.testing.assert_allclose(w, a["w"], rtol=0, atol=2e-6)
    assert np.abs(g_xyz - a["g_xyz"]).max() <= 1e-5 * np.abs(a["g_xyz"]).max() * 5       # fp32 near cell faces: d/dx of a product of differences
    assert np.abs(g_grid - a["g_grid"]).max() <= 1e-5 * np.abs(a["g_grid"]).max()


def test_skin_weights_realistic_grid_vs_oracle(built_lib):
    """Reference-sized problem: 116 x 142 x 196 x 21 grid (res 128, ratios 1.1 / 0.9 / 0.65) and 100k points."""
    rng = np.random.default_rng(0)
    D, H, W, C = 196, 142, 116, 21
    grid = rng.uniform(0, 1, (D, H, W, C)).astype(np.float32) * (rng.uniform(0, 1, (D, H, W, C)) < 0.15)
    grid[..., -1] += 0.05
    center, scale = np.array([0.0, 0.01, 0.1], np.float32), np.array([[0.14, 0.12, 0.09]], np.float32)
    xyz = (rng.uniform(-0.97, 0.97, (100_000, 3)) * scale + center).astype(np.float32)
    gout = rng.normal(0, 1, (xyz.shape[0], C)).astype(np.float32)
    w, g_xyz, _ = run(xyz, center, scale, grid, gout, grid_grad=False)
    w_ref, _ = skin_ref.skin_weights(xyz, center, scale, grid)
    g_ref, _ = skin_ref.skin_weights_backward(xyz, center, scale, grid, gout)
    # fp32 cell coordinates reach 195 (ulp 1.5e-5) and carry the rounding of the division: up to ~1e-4 on single weights
    err_w = np.abs(w - w_ref)
    assert np.percentile(err_w, 99.99) <= 3e-5 and err_w.max() <= 2e-4
    assert np.abs(w.sum(-1) - 1).max() <= 1e-5
    err = np.abs(g_xyz - g_ref)
    assert np.percentile(err, 99.9) <= 1e-4 * np.abs(g_ref).max() and err.max() <= 2e-3 * np.abs(g_ref).max()
