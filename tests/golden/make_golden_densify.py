"""Golden vectors for densification / pruning and the optimizer-state surgery FROM THE REFERENCE'S OWN PYTHON:
    python tests/golden/make_golden_densify.py      -> tests/golden/densify_golden.npz
Runs src/models/gaussian.py GaussianModel.training_setup, add_densification_stats, densify_and_prune (clone + split + prune,
:148-338) and reset_opacity on a CPU model with an Adam state that has seen two steps; records parameters, both Adam
moments, skin weights and statistics before and after.  torch.manual_seed fixes the samples of densify_and_split.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import as R  # noqa: E402

R.install()
import torch  # noqa: E402

import src.models.gaussian as ref_gm  # noqa: E402

torch.set_num_threads(4)
NAMES = ["xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"]
ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity", "scaling": "_scaling", "rotation": "_rotation"}


def make_model(N, seed):
    g = torch.Generator().manual_seed(seed)
    m = object.__new__(ref_gm.GaussianModel)
    torch.nn.Module.__init__(m)
    m.opts = types.SimpleNamespace(isotropic_scaling=False, sh_degree=3, percent_dense=0.01, position_lr_init=0.00016, position_lr_final=0.0000016,
                                   position_lr_delay_mult=0.01, position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.05,
                                   scaling_lr=0.005, rotation_lr=0.001)
    m.spatial_lr_scale = 0.2
    m.setup_functions()
    m._xyz = torch.nn.Parameter(torch.randn(N, 3, generator=g) * 0.05)
    m._scaling = torch.nn.Parameter(torch.log(torch.rand(N, 3, generator=g) * 0.004 + 0.0002))
    m._rotation = torch.nn.Parameter(torch.randn(N, 4, generator=g))
    m._opacity = torch.nn.Parameter(torch.randn(N, 1, generator=g) * 3.0)
    m._features_dc = torch.nn.Parameter(torch.randn(N, 1, 3, generator=g))
    m._features_rest = torch.nn.Parameter(torch.randn(N, 15, 3, generator=g) * 0.1)
    w = torch.rand(N, 21, generator=g) * (torch.rand(N, 21, generator=g) < 0.2)
    w[:, -1] += 0.05
    m._skin_weights = w / w.sum(1, keepdim=True)
    m.max_radii2D = torch.zeros(N)
    m.training_setup()
    for _ in range(2):                                           # give Adam non-trivial moments
        for name in NAMES:
            p = getattr(m, ATTR[name])
            p.grad = torch.randn(p.shape, generator=g) * 1e-3
        m.optimizer.step()
    return m, g


def snapshot(m, prefix, out):
    for name in NAMES:
        p = getattr(m, ATTR[name])
        out[f"{prefix}_{name}"] = p.detach().numpy().copy()
        st = m.optimizer.state[p]
        out[f"{prefix}_{name}_exp_avg"] = st["exp_avg"].numpy().copy()
        out[f"{prefix}_{name}_exp_avg_sq"] = st["exp_avg_sq"].numpy().copy()
    out[f"{prefix}_skin"] = m._skin_weights.numpy().copy()
    out[f"{prefix}_accum"] = m.xyz_gradient_accum.numpy().copy()
    out[f"{prefix}_denom"] = m.denom.numpy().copy()
    out[f"{prefix}_max_radii2D"] = m.max_radii2D.numpy().copy()


def main():
    out = {}
    N = 300
    m, g = make_model(N, 0)
    snapshot(m, "before", out)
    # three views of densification statistics (gaussian.py:335-338, gaussian_utils.py:461-473)
    for v in range(3):
        vs = types.SimpleNamespace(grad=torch.randn(N, 3, generator=g) * 4e-4)
        filt = torch.rand(N, generator=g) < 0.7
        radii = (torch.rand(N, generator=g) * 40).int() * filt
        m.max_radii2D[filt] = torch.max(m.max_radii2D[filt], radii[filt].float())
        m.add_densification_stats(vs, filt)
        out[f"view{v}_grad"], out[f"view{v}_filter"], out[f"view{v}_radii"] = vs.grad.numpy().copy(), filt.numpy().copy(), radii.numpy().copy()
    snapshot(m, "stats", out)
    torch.manual_seed(1234)
    m.densify_and_prune(0.0002, 0.005, 0.25, 20)
    snapshot(m, "after", out)
    m.reset_opacity()
    snapshot(m, "reset", out)
    out["args"] = np.array([0.0002, 0.005, 0.25, 20.0, 1234.0])
    np.savez_compressed(os.path.join(HERE, "densify_golden.npz"), **out)
    print("N before", N, "after", out["after_xyz"].shape[0])


if __name__ == "__main__":
    main()
