"""The SH -> RGB helper of manus_b200.render (plain torch, reference semantics) against the pinned pose oracle."""
import types

import torch

from manus_b200.render import calculate_colors_from_sh
from oracle import pose_ref


def test_calculate_colors_from_sh_matches_oracle():
    g = torch.Generator().manual_seed(0)
    n = 500
    feats = torch.randn(n, 16, 3, generator=g) * 0.3
    cano = torch.randn(n, 3, generator=g) * 0.05
    tf = torch.eye(4).repeat(n, 1, 1)
    tf[:, :3, :3] += 0.1 * torch.randn(n, 3, 3, generator=g)
    tf[:, :3, 3] = 0.03 * torch.randn(n, 3, generator=g)
    posed = torch.einsum("nij,nj->ni", tf, torch.cat([cano, torch.ones(n, 1)], 1))[:, :3]
    cam = types.SimpleNamespace(camera_center=torch.tensor([0.2, -0.1, 1.3]))
    for deg in (0, 1, 2, 3):
        for t in (tf, None):
            got = calculate_colors_from_sh(posed, feats, cano, cam, deg, t)
            ref = pose_ref.calculate_colors_from_sh(posed, feats, cano, cam.camera_center, deg, t)
            assert torch.allclose(got, ref, atol=1e-6), (deg, t is None)
