"""``render_gaussians``: the caller of the rasterizer boundary, same signature and return dict as
/root/reference/src/utils/gaussian_utils.py:349-428, plus the fused entry ``render_fused`` that starts from the raw
GaussianModel parameters (one pose kernel + the rasterizer; nothing else runs per frame).
"""
from __future__ import annotations

import math

import torch

from .pose import pose_gaussians
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer


def _scalar(v) -> float:
    return float(v.reshape(-1)[0]) if torch.is_tensor(v) else float(v)


def _settings(camera, bg_color, sh_degree, device) -> GaussianRasterizationSettings:
    tan_dev = getattr(camera, "tanfov_dev", None)     # optional CUDA tensor (tanfovx, tanfovy): intrinsics stay on the device
    return GaussianRasterizationSettings(
        image_height=int(_scalar(camera.height)), image_width=int(_scalar(camera.width)),
        tanfovx=tan_dev[0:1] if tan_dev is not None else math.tan(_scalar(camera.fovx) * 0.5),
        tanfovy=tan_dev[1:2] if tan_dev is not None else math.tan(_scalar(camera.fovy) * 0.5),
        bg=bg_color, scale_modifier=1,
        viewmatrix=torch.as_tensor(camera.world_view_transform).to(device),
        projmatrix=torch.as_tensor(camera.full_proj_transform).to(device),
        sh_degree=sh_degree, campos=torch.as_tensor(camera.camera_center).to(device), prefiltered=False, debug=False)


def calculate_colors_from_sh(posed_means, cano_features, cano_means, camera, sh_degree, tf):
    """gaussian_utils.py:431-449 for callers that hold the materialised per-Gaussian ``tf`` [N,4,4]: evaluated by the
    fused pose kernel with tf given as N one-bone skinning (weights = 1)."""
    raise NotImplementedError("use render_fused / pose_gaussians (the fused path never materialises tf)")


_ZEROS = {}


def _screenspace_leaf(posed_means):
    """A [N,3] zero tensor that receives the screen-space gradient, without a fill kernel per frame: the rasterizer never
    reads means2D, so a cached zero buffer is re-used and only a fresh autograd leaf is made on top of it."""
    key = (posed_means.device, tuple(posed_means.shape), posed_means.dtype)
    z = _ZEROS.get(key)
    if z is None:
        _ZEROS.clear()
        z = _ZEROS[key] = torch.zeros_like(posed_means, requires_grad=False)
    return z.detach().requires_grad_(True)


def render_gaussians(posed_means, posed_cov, cano_means, cano_features, cano_opacity, camera, bg_color, colors_precomp=None,
                     sh_degree=3, tf=None, device=torch.device("cuda"), _cached_screenspace=False):
    """gaussian_utils.py:349-428.  ``colors_precomp`` must be given (MANUS computes it with calculate_colors_from_sh
    before the call, :401-404; in this package colours come out of ``pose_gaussians``)."""
    if _cached_screenspace:
        screenspace_points = _screenspace_leaf(posed_means)          # a leaf: .grad is populated without retain_grad()
    else:
        screenspace_points = torch.zeros_like(posed_means, dtype=posed_means.dtype, requires_grad=True, device=device) + 0
        try:
            screenspace_points.retain_grad()
        except Exception:
            pass
    if colors_precomp is None:
        raise ValueError("render_gaussians: pass colors_precomp (from manus_b200.pose_gaussians)")
    rasterizer = GaussianRasterizer(raster_settings=_settings(camera, bg_color, sh_degree, device))
    rendered_image, radii = rasterizer(means3D=posed_means, means2D=screenspace_points, shs=None, colors_precomp=colors_precomp,
                                       opacities=cano_opacity, scales=None, rotations=None, cov3D_precomp=posed_cov)
    rendered_image = torch.permute(rendered_image, (1, 2, 0))
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii}


def render_fused(params, skin_wts, bone_tf, camera, bg_color, sh_degree=3, isotropic=False, num_skinned=None, grad_sink=None):
    """params: (xyz, log_scale, quat, opacity_logit, f_dc, f_rest) -- the six nn.Parameters of GaussianModel.
    Returns the render_gaussians dict plus the posed quantities (what TrainingModule.forward returns)."""
    xyz, log_scale, quat, opacity_logit, f_dc, f_rest = params
    device = xyz.device
    campos = torch.as_tensor(camera.camera_center).to(device)
    posed_xyz, posed_cov, colors, opacity = pose_gaussians(xyz, log_scale, quat, opacity_logit, f_dc, f_rest, skin_wts, bone_tf,
                                                          campos, sh_degree, isotropic, num_skinned, grad_sink)
    out = render_gaussians(posed_xyz, posed_cov, xyz, None, opacity, camera, bg_color, colors_precomp=colors,
                           sh_degree=sh_degree, device=device, _cached_screenspace=True)
    out.update(posed_xyz=posed_xyz, posed_cov=posed_cov, colors=colors, cano_opacity=opacity)
    return out
