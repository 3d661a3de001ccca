"""Parity against the reference's OWN CUDA rasterizer (graphdeco-inria/diff-gaussian-rasterization, the module MANUS imports at
src/utils/gaussian_utils.py:18-21).  Its sources are not part of /root/reference (setup_env.sh:4-13 clones them at install time)
and this image has no network, so the test is skipped unless somebody provisions the built extension under baseline/_ref; the
day that happens it unpins R2 / R3: image, radii and all five gradients of the existing small scenes, north_star's 1e-5."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _upstream():
    import bench_extras

    mod = bench_extras.import_upstream()
    if mod is None:
        pytest.skip("upstream diff_gaussian_rasterization is not installed under baseline/_ref (absent from the reference checkout, no network)")
    return mod


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_image_radii_and_gradients_match_upstream(built_lib, seed):
    from bench_extras import render_like_manus
    from helpers import grad_close
    from test_gpu_raster import small_scene

    up = _upstream()
    sys.path.insert(0, os.path.join(ROOT, "shims"))
    import diff_gaussian_rasterization as ours

    dev = torch.device("cuda", 0)
    sc, cam, ps = small_scene(seed)
    G = torch.rand(cam.height, cam.width, 3, generator=torch.Generator().manual_seed(7)).to(dev)
    bg = torch.ones(3, device=dev)
    res = {}
    for name, mod in (("upstream", up), ("ours", ours)):
        leaves = {k: torch.tensor(np.ascontiguousarray(ps[k]), device=dev).requires_grad_(True) for k in ("means3D", "cov3D", "opacity", "colors")}
        img, radii, screen = render_like_manus(mod, leaves["means3D"], leaves["cov3D"], leaves["opacity"], leaves["colors"], cam, bg, dev)
        (img * G).sum().backward()
        res[name] = (img.detach().cpu().numpy(), radii.cpu().numpy(), {k: v.grad.cpu().numpy() for k, v in leaves.items()},
                     screen.grad.cpu().numpy())
    np.testing.assert_array_equal(res["ours"][1], res["upstream"][1])
    err = np.abs(res["ours"][0] - res["upstream"][0])
    assert (err > 1e-5).mean() <= 1e-3 and err.max() <= 2e-2, (float(err.max()), float((err > 1e-5).mean()))   # gate-fragile pixels aside
    for k in res["ours"][2]:
        ok, e, s = grad_close(res["ours"][2][k], res["upstream"][2][k])
        assert ok, (k, e, s)
    ok, e, s = grad_close(res["ours"][3], res["upstream"][3])
    assert ok, ("means2D", e, s)
