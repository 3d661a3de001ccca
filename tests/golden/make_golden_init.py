"""Golden vectors for the parameter initialisation FROM THE REFERENCE'S OWN PYTHON (needs /root/reference):
    python tests/golden/make_golden_init.py          -> tests/golden/init_golden.npz
Runs ``GaussianModel.initialize_parameters`` (src/models/gaussian.py:99-127) on CPU.  Its ``distCUDA2`` (:110) is a CUDA
extension that is not part of the reference tree; the exact float64 k-d-tree statistic (oracle/knn_ref.dist2_knn3) stands
in for it here -- everything downstream of that call (clamp, log sqrt, RGB2SH, inverse_sigmoid, layouts) is the reference's code."""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _ref_import as R  # noqa: E402

R.install()
import torch  # noqa: E402

import src.models.gaussian as ref_gm  # noqa: E402
from oracle import knn_ref  # noqa: E402

ref_gm.distCUDA2 = lambda pts: torch.from_numpy(knn_ref.dist2_knn3(pts.cpu().numpy()))
ref_gm.cprint = lambda *a, **k: None


def run(points, colors, isotropic):
    m = object.__new__(ref_gm.GaussianModel)
    torch.nn.Module.__init__(m)
    m.opts = types.SimpleNamespace(isotropic_scaling=isotropic, sh_degree=3)
    m.max_sh_degree = 3
    m.setup_functions()
    m.initialize_parameters(points, colors, device=torch.device("cpu"))
    return {"xyz": m._xyz, "f_dc": m._features_dc, "f_rest": m._features_rest, "log_scale": m._scaling, "quat": m._rotation,
            "opacity_logit": m._opacity}


def main():
    rng = np.random.default_rng(3)
    n = 700
    pts = rng.normal(0, 0.04, (n, 3)).astype(np.float32)
    pts[5] = pts[2]; pts[9] = pts[2]; pts[11] = pts[2]           # four coincident points: dist2 = 0 -> the 1e-7 clamp
    cols = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    out = {"points": pts, "colors": cols, "dist2": knn_ref.dist2_knn3(pts)}
    for iso in (False, True):
        for k, v in run(pts, cols, iso).items():
            out[f"{'iso' if iso else 'aniso'}_{k}"] = v.detach().numpy()
    path = os.path.join(HERE, "init_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
