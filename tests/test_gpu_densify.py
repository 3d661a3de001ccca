"""Densify / prune / opacity reset with the optimizer-state surgery (src/models/gaussian.py:148-338) ON THE DEVICE, inside a
training loop through the CUDA path.  The CPU suite (tests/test_densify.py) pins the arithmetic bit-exactly to goldens produced by
the reference's own GaussianModel; torch.normal draws differ between CPU and CUDA generators, so here the same step is checked
through what does not depend on the draws: which Gaussians are selected (against the reference's selection rule restated on the
statistics the pose backward kernel accumulated), the bookkeeping of parameters / Adam moments / skin weights / statistics, and
that renderer and captured steps follow the rebuilt buffers (or refuse to run on stale ones)."""
import numpy as np
import pytest
import torch

from helpers import zoom_camera
from manus_b200 import synth

pytestmark = pytest.mark.gpu


def test_densify_and_prune_on_device_after_real_steps(built_lib):
    from manus_b200 import rasterizer as rz
    from manus_b200.densify import GaussianState
    from manus_b200.dist import GraphedStep, SceneRenderer, pack_camera
    from manus_b200.optim import FlatAdam

    dev = torch.device("cuda", 0)
    W, H = 256, 144
    scene = synth.make_composite(6000, seed=4)
    r = SceneRenderer(scene, dev, W, H)
    for view in (2, 7):
        cam = zoom_camera(view, W, H, 1.3)
        r._cams[view] = (cam, torch.from_numpy(pack_camera(cam)), torch.from_numpy(synth.posed_bones(view).reshape(-1).astype("float32")))
    opt = FlatAdam(r.flat, dict(xyz=1.6e-4, f_dc=2.5e-3, f_rest=2.5e-3 / 20, opacity=5e-2, scaling=5e-3, rotation=1e-3))
    skin_full = torch.cat([r.skin, torch.zeros(scene.n - scene.n_hand, r.skin.shape[1], device=dev)], 0)   # one row per Gaussian for the surgery
    gs = GaussianState(r.flat, opt, skin_full)
    G = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(3)).to(dev)
    rz.set_capacity_mode("exact")
    try:
        dmax = 0
        for view in (2, 7):
            _, c, b = r.view_inputs_host(view)
            r.render(view, cam_dev=c.to(dev), bones_dev=b.to(dev))
            dmax = max(dmax, rz.check_overflow())
        rz.set_capacity_mode("reserve", margin=1.3)
        rz.reserve_capacity(dev.index, scene.n, H, W, dmax)
        step = GraphedStep(r, lambda image, target: (image * target).sum(), G, view=2,
                           stats=(gs.xyz_gradient_accum, gs.denom, gs.max_radii2D))
        for it in range(6):                                   # a few real steps: statistics from the kernel, Adam moments non-zero
            _, c, b = r.view_inputs_host((2, 7)[it % 2])
            step.set_inputs(c.to(dev), b.to(dev))
            step.replay()
            r.flat.grad.mul_(1e-6)
            opt.step()
        torch.cuda.synchronize()
        assert float(gs.denom.max()) == 6.0 and float(gs.xyz_gradient_accum.max()) > 0
        n0 = gs.n
        op0 = torch.sigmoid(r.flat.params["opacity_logit"]).squeeze(1).clone()
        grads = (gs.xyz_gradient_accum / gs.denom).nan_to_num(0.0)
        max_grad = float(torch.quantile(grads[grads > 0], 0.9))
        extent, min_opacity = 0.3, 0.02
        # the reference's selection rules (gaussian.py:249-304) restated on the same statistics
        big = gs.get_scaling.max(dim=1).values > gs.percent_dense * extent
        sel = grads.squeeze(1) >= max_grad
        clone, split = sel & ~big, sel & big
        n_clone, n_split = int(clone.sum()), int(split.sum())
        assert n_clone + n_split > 0
        low = op0 < min_opacity
        pruned = int(low[~split].sum()) + int(low[clone].sum()) + 2 * int(low[split].sum())
        quat_kept = r.flat.params["quat"][~split & ~low].clone()
        gs.densify_and_prune(max_grad, min_opacity, extent, max_screen_size=None)
        torch.cuda.synchronize()
        assert gs.n == n0 + n_clone + n_split - pruned          # + clones, + 2 children - 1 parent per split, - faint ones
        assert gs.n == gs.flat.n == gs.skin_wts.shape[0] == gs.max_radii2D.shape[0] and gs.flat.grad.numel() == gs.flat.data.numel()
        assert torch.equal(gs.flat.params["quat"][: quat_kept.shape[0]], quat_kept)                  # survivors keep their rows, in order
        assert float(gs.xyz_gradient_accum.abs().max()) == 0 and float(gs.denom.abs().max()) == 0      # statistics reset (gaussian.py:245-247)
        assert gs.opt.exp_avg.numel() == gs.flat.data.numel() and bool(torch.isfinite(gs.flat.data).all())
        assert float(gs.opt.exp_avg.abs().max()) > 0           # moments of the surviving Gaussians were carried over
        # the captured step holds the addresses of the old buffers: after the renderer follows the new ones it must refuse to replay
        r.rebind(gs.flat, None)
        assert r.flat.n == gs.n
        with pytest.raises(RuntimeError, match="replaced"):
            step.replay()
        gs.reset_opacity()
        assert float(torch.sigmoid(gs.flat.params["opacity_logit"]).max()) <= 0.0100001
    finally:
        rz.set_capacity_mode("exact")
