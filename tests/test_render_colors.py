"""The SH -> RGB entry of manus_b200.render (calculate_colors_from_sh, gaussian_utils.py:431-449) against the pinned pose oracle,
all SH degrees, with and without a materialised tf; and its refusal to compute anything on the CPU."""
import types

import pytest
import torch

from manus_b200.render import calculate_colors_from_sh
from oracle import pose_ref


def _inputs(device):
    g = torch.Generator().manual_seed(0)
    n = 500
    feats = torch.randn(n, 16, 3, generator=g) * 0.3
    cano = torch.randn(n, 3, generator=g) * 0.05
    tf = torch.eye(4).repeat(n, 1, 1)
    tf[:, :3, :3] += 0.1 * torch.randn(n, 3, 3, generator=g)
    tf[:, :3, 3] = 0.03 * torch.randn(n, 3, generator=g)
    posed = torch.einsum("nij,nj->ni", tf, torch.cat([cano, torch.ones(n, 1)], 1))[:, :3]
    cam = types.SimpleNamespace(camera_center=torch.tensor([0.2, -0.1, 1.3]))
    return [t.to(device) for t in (feats, cano, tf, posed)], cam


def test_calculate_colors_from_sh_has_no_cpu_path():
    from manus_b200 import _lib

    (feats, cano, tf, posed), cam = _inputs("cpu")
    with pytest.raises(_lib.ManusB200Error, match="no CPU path"):
        calculate_colors_from_sh(posed, feats, cano, cam, 3, tf)


@pytest.mark.gpu
def test_calculate_colors_from_sh_matches_oracle(built_lib):
    (feats, cano, tf, posed), cam = _inputs("cuda")
    for deg in (0, 1, 2, 3):
        for t in (tf, None):
            got = calculate_colors_from_sh(posed, feats, cano, cam, deg, t)
            ref = pose_ref.calculate_colors_from_sh(posed, feats, cano, cam.camera_center.to("cuda"), deg, t)
            assert torch.allclose(got, ref, atol=2e-6), (deg, t is None)
