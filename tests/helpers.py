"""Shared builders for the parity tests (seeded scenes, oracle / product drivers, comparison rules)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from manus_b200 import synth  # noqa: E402
from manus_b200.cameras import opengl_camera  # noqa: E402
from oracle import pose_ref  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# Tolerances (north_star: 1e-5 abs on images and gradients for identical inputs).
IMG_ATOL = 1e-5
# Gradients of sum(image * G) are O(1e2..1e5); 1e-5 is applied relative to the largest magnitude of each gradient
# tensor (with an absolute floor of 1e-5).  fp32-oracle vs fp64-oracle noise is ~2e-6 in the same norm.
GRAD_RTOL = 1e-5


def grad_close(got: np.ndarray, ref: np.ndarray, rtol: float = GRAD_RTOL):
    scale = max(1.0, float(np.abs(ref).max()) if ref.size else 1.0)
    err = float(np.abs(got - ref).max()) if ref.size else 0.0
    return err <= rtol * scale, err, scale


def zoom_camera(view: int, W: int, H: int, zoom: float = 1.0):
    """A shipped camera rescaled to W x H, optionally zoomed in so that the hand fills a small test image."""
    fx = synth.fixtures()
    f_x, f_y = fx["cam_intrs"][view % 51][:2]
    return opengl_camera(f_x * W / 1920.0 * zoom, f_y * H / 1080.0 * zoom, fx["cam_extrs"][view % 51], W, H)


def posed_scene(scene, view: int, cam, scale_boost: float = 0.0, opacity_boost: float = 0.0):
    """Run the (pinned) pose oracle on a synth.Scene -> numpy rasterizer inputs."""
    t = lambda a: None if a is None else torch.tensor(a)
    if scene.n_hand > 0:
        tfs = pose_ref.bone_transforms(t(synth.posed_bones(view)), t(scene.bones_rest), True)
    outs = []
    nh = scene.n_hand
    sl = lambda a, lo, hi: t(a[lo:hi])
    parts = []
    if nh > 0:
        parts.append((0, nh, t(scene.skin_wts), tfs))
    if nh < scene.n:
        parts.append((nh, scene.n, None, None))
    for lo, hi, w, tf in parts:
        outs.append(pose_ref.pose_gaussians_ref(sl(scene.xyz, lo, hi), sl(scene.log_scale, lo, hi) + scale_boost, sl(scene.quat, lo, hi),
                                                sl(scene.opacity_logit, lo, hi) + opacity_boost, sl(scene.f_dc, lo, hi),
                                                sl(scene.f_rest, lo, hi), w, tf, t(cam.camera_center)))
    cat = lambda k: torch.cat([o[k] for o in outs], 0).numpy()
    return dict(means3D=cat(0), cov3D=cat(1), colors=cat(2), opacity=cat(3))


def cam_args(cam, bg=(1.0, 1.0, 1.0)):
    return dict(viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, campos=cam.camera_center,
                tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, W=cam.width, H=cam.height, bg=np.asarray(bg, np.float32))


def settings_from(cam, bg, device, sh_degree=3, debug=False):
    from manus_b200.rasterizer import GaussianRasterizationSettings

    return GaussianRasterizationSettings(
        image_height=cam.height, image_width=cam.width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=torch.tensor(np.asarray(bg, np.float32), device=device), scale_modifier=1.0,
        viewmatrix=torch.tensor(cam.world_view_transform, device=device),
        projmatrix=torch.tensor(cam.full_proj_transform, device=device), sh_degree=sh_degree,
        campos=torch.tensor(cam.camera_center, device=device), prefiltered=False, debug=debug)


def fragile_pixels(ref64_img: np.ndarray, ref32_img: np.ndarray, tol: float = IMG_ATOL) -> np.ndarray:
    """Pixels where the fp32 and fp64 oracles themselves disagree by more than the tolerance: a hard gate
    (alpha >= 1/255, T >= 1e-4, power <= 0) flipped under rounding.  [H,W] bool."""
    return (np.abs(ref64_img - ref32_img) > tol).any(0)


def small_scene_inputs(seed: int, N: int = 3000, W: int = 160, H: int = 120, zoom: float = 1.6, device="cuda"):
    """(camera, {means3D, cov3D, colors, opacity} as device tensors) of a small posed hand scene."""
    sc = synth.make_hand(N, seed=seed)
    cam = zoom_camera(seed * 7 % 51, W, H, zoom)
    ps = posed_scene(sc, 3 * seed + 1, cam, 0.3, 1.0)
    return cam, {k: torch.tensor(np.ascontiguousarray(v), device=device) for k, v in ps.items()}
