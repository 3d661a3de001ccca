#!/usr/bin/env python
"""One small pass through every kernel family of the library, sized for compute-sanitizer (the
reference has no race detection; this build runs memcheck / racecheck on small scenes):

    compute-sanitizer --tool memcheck  --error-exitcode 9 python tools/sanitize_scene.py
    compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_scene.py
    compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_scene.py

Kernels launched: pose forward / backward (overwrite and TMA reduce-add), preprocess, radix histogram / passes (depth and
tile sorts), instance emission, tile ranges / order / segment items, blend forward / backward, preprocess backward (all four
colour / covariance modes), photometric loss, fused Adam, SH-gradient rebuild, skin-weight lookup forward / backward, 3-NN
distances, nearest reference point."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import zoom_camera  # noqa: E402
from manus_b200 import synth  # noqa: E402
from manus_b200.dist import SceneRenderer, pack_camera  # noqa: E402
from manus_b200.knn import distCUDA2, get_contact_dist  # noqa: E402
from manus_b200.losses import photometric_loss  # noqa: E402
from manus_b200.optim import FlatAdam  # noqa: E402
from manus_b200.pose import sh_grad_from_views  # noqa: E402
from manus_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402
from manus_b200.skinning import skinning_weights_from_voxel_grid  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3001        # ragged last tiles everywhere
W, H = 176, 100                                             # ragged tile grid (11 x 6.25 tiles)
dev = torch.device("cuda", 0)
scene = synth.make_composite(N, seed=2)
r = SceneRenderer(scene, dev, W, H)
for view in (3, 9):
    cam = zoom_camera(view, W, H, 1.4)
    r._cams[view] = (cam, torch.from_numpy(pack_camera(cam)), torch.from_numpy(synth.posed_bones(view).reshape(-1).astype("float32")))
G = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(1)).to(dev)

# fused path: pose -> raster -> photometric loss -> raster backward -> pose backward (overwrite, then accumulate)
for k, view in enumerate((3, 9)):
    out = r.render(view, sink=r.flat.grads, accumulate=k > 0)
    photometric_loss(out["render"], G, 0.8, 0.2).backward()
g_two_views = r.flat.grad.clone()
assert torch.isfinite(g_two_views).all() and float(g_two_views.abs().max()) > 0

# optimizer step on the flat buffers
opt = FlatAdam(r.flat, dict(xyz=1.6e-4, f_dc=2.5e-3, f_rest=2.5e-3 / 20, opacity=5e-2, scaling=5e-3, rotation=1e-3))
opt.step()

# SH-gradient rebuild from the views' DC gradients (the compact exchange's local step)
cam = r._cams[9][0]
bone_all = r._bone_tf[None].contiguous()
campos_all = torch.tensor(np.asarray(cam.camera_center, np.float32), device=dev)[None]
gfdc_all = r.flat.grads["f_dc"].reshape(1, -1, 3).clone()
sh_grad_from_views(r.flat.params["xyz"], r.skin, r.n_hand, 3, r.flat.sh_coeffs, bone_all, campos_all, gfdc_all, r.flat.grads["f_dc"],
                   r.flat.grads["f_rest"])

# the rasterizer in upstream's other input mode: SH colours and scale / rotation covariances evaluated in the raster kernels
t = lambda a: torch.tensor(np.ascontiguousarray(a), device=dev)
means = t(scene.xyz).requires_grad_(True)
shs = t(np.concatenate([scene.f_dc, scene.f_rest], 1)).requires_grad_(True)
scales, rots = t(np.exp(scene.log_scale + 0.3)).requires_grad_(True), t(scene.quat).requires_grad_(True)
op = torch.sigmoid(t(scene.opacity_logit) + 1.0).requires_grad_(True)
rs = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, torch.ones(3, device=dev), 1.0, t(cam.world_view_transform),
                                   t(cam.full_proj_transform), 3, t(cam.camera_center), False, False)
m2 = torch.zeros_like(means, requires_grad=True)
img, radii = GaussianRasterizer(rs)(means3D=means, means2D=m2, opacities=op, shs=shs, scales=scales, rotations=rots)
(img.permute(1, 2, 0) * G).sum().backward()
assert torch.isfinite(shs.grad).all() and torch.isfinite(rots.grad).all()

# skin-weight lookup (trilinear gather + normalisation), forward and backward
grid = torch.rand(12, 14, 16, 21, generator=torch.Generator().manual_seed(4)).to(dev).requires_grad_(True)
pts = (torch.rand(2003, 3, generator=torch.Generator().manual_seed(5)).to(dev) - 0.5).requires_grad_(True)
w = skinning_weights_from_voxel_grid(pts, torch.zeros(3, device=dev), torch.full((3,), 0.6, device=dev), grid)
(w * torch.arange(21, device=dev)).sum().backward()

# 3-NN mean squared distance and nearest reference point
d2 = distCUDA2(t(scene.xyz))
dist = get_contact_dist(t(scene.xyz[: scene.n_hand]), t(scene.xyz[scene.n_hand:]))
torch.cuda.synchronize()
assert torch.isfinite(d2).all() and float(d2.min()) >= 0
print(f"sanitize scene ok: N={N} {W}x{H} |grad|max={float(g_two_views.abs().max()):.3e} mean d2={float(d2.mean()):.3e}")
