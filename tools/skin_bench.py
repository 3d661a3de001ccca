#!/usr/bin/env python
"""Micro-benchmark of the skin-weight lookup at the reference's grid size (116 x 142 x 196 x 21) and 300k points."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manus_b200 import _lib  # noqa: E402
from manus_b200.skinning import skinning_weights_from_voxel_grid  # noqa: E402

D, H, W, C = 196, 142, 116, 21
grid = torch.rand(D, H, W, C, device="cuda")
center = torch.tensor([0., 0.01, 0.1]).cuda()
scale = torch.tensor([[0.14, 0.12, 0.09]]).cuda()
spread = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
xyz = ((torch.rand(300000, 3, device="cuda") * 2 - 1) * spread * scale + center).requires_grad_(True)
_lib.profile_enable(True)
_lib.profile_report()
for _ in range(6):
    w = skinning_weights_from_voxel_grid(xyz, center, scale, grid)
    w.sum().backward()
print({k: (n, round(ms / n * 1e3, 1)) for k, (n, ms) in _lib.profile_report().items()})
