"""Golden vectors for the per-frame skin-weight lookup FROM THE REFERENCE'S OWN PYTHON (needs /root/reference):
    python tests/golden/make_golden_skin.py          -> tests/golden/skin_golden.npz
Runs src/utils/gaussian_utils.py:167-196 ``skinning_weights_from_voxel_grid`` (grid_sample, align_corners=True, zero
padding, row normalisation) as ``HandGaussianModel.get_skin_weights`` calls it (src/models/hand_gaussian.py:65-76), and torch
autograd through it for d/d xyz and d/d grid_weights.  The grid mimics ``build_voxel_grid`` (src/datasets/brics_dynamic.py:99-144):
[D,H,W,21] non-negative weights with ~3 non-zeros per cell, per-axis scale, centre offset.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import as R  # noqa: E402

R.install()
import torch  # noqa: E402

import src.utils.gaussian_utils as ref_gu  # noqa: E402

torch.set_num_threads(4)


def case(D, H, W, C, N, seed):
    g = torch.Generator().manual_seed(seed)
    dense = torch.rand(D, H, W, C, generator=g)
    keep = torch.rand(D, H, W, C, generator=g) < 0.18                       # sparse rows like MANO weights
    keep[..., -1] |= ~keep.any(-1)                                          # at least the background bone
    grid = (dense * keep).float()
    grid = grid / grid.sum(-1, keepdim=True)
    center = torch.tensor([0.01, -0.02, 0.1])
    scale = torch.tensor([[0.13 * 1.1, 0.13 * 0.9, 0.13 * 0.65]])           # build_voxel_grid: scale * (z,y,x) ratios
    xyz = (torch.rand(N, 3, generator=g) * 2 - 1) * scale * 1.08 + center   # a few points fall outside the grid
    xyz[0] = center + scale[0] * torch.tensor([1.0, -1.0, 1.0])             # exactly on a corner
    xyz[1] = center                                                        # exactly in the middle
    xyz = xyz.requires_grad_(True)
    grid_p = grid.clone().requires_grad_(True)
    w = ref_gu.skinning_weights_from_voxel_grid(xyz, center, scale, grid_p)
    gout = torch.randn(N, C, generator=g)
    ok = torch.isfinite(w).all(-1)
    (w[ok] * gout[ok]).sum().backward()
    return dict(grid=grid.numpy(), center=center.numpy(), scale=scale.numpy(), xyz=xyz.detach().numpy(), w=w.detach().numpy(),
                gout=gout.numpy(), finite=ok.numpy(), g_xyz=xyz.grad.numpy(), g_grid=grid_p.grad.numpy())


def main():
    out = {}
    for name, args in {"a": (9, 11, 13, 21, 400, 0), "b": (6, 5, 7, 20, 257, 1), "c": (2, 2, 2, 3, 64, 2)}.items():
        for k, v in case(*args).items():
            out[f"{name}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "skin_golden.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("_w")}, {k: int((~v).sum()) for k, v in out.items() if k.endswith("finite")})


if __name__ == "__main__":
    main()
