#!/usr/bin/env python
"""Kernel time of the fused photometric loss at 1080p (pred in the rasterizer's planar layout)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manus_b200 import _lib  # noqa: E402
from manus_b200.losses import photometric_loss  # noqa: E402

g = torch.rand(1080, 1920, 3, device="cuda")
chw = (g + 0.1 * torch.randn_like(g)).permute(2, 0, 1).contiguous().requires_grad_(True)
_lib.profile_enable(True)
_lib.profile_report()
for _ in range(10):
    photometric_loss(chw.permute(1, 2, 0), g).backward()
print({k: (n, round(ms / n * 1e3, 1)) for k, (n, ms) in _lib.profile_report().items()})
