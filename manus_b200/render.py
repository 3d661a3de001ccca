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


_SH_C0 = 0.28209479177387814
_SH_C1 = 0.4886025119029199
_SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435)


def eval_sh(deg: int, sh: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """src/utils/sh_utils.py:57-120 for degrees 0..3: sh [N,3,K], unit dirs [N,3] -> [N,3] (plain torch, any device)."""
    result = _SH_C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        result = result - _SH_C1 * y * sh[..., 1] + _SH_C1 * z * sh[..., 2] - _SH_C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            result = (result + _SH_C2[0] * xy * sh[..., 4] + _SH_C2[1] * yz * sh[..., 5] + _SH_C2[2] * (2.0 * zz - xx - yy) * sh[..., 6]
                      + _SH_C2[3] * xz * sh[..., 7] + _SH_C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                result = (result + _SH_C3[0] * y * (3 * xx - yy) * sh[..., 9] + _SH_C3[1] * xy * z * sh[..., 10]
                          + _SH_C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + _SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
                          + _SH_C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + _SH_C3[5] * z * (xx - yy) * sh[..., 14]
                          + _SH_C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return result


def calculate_colors_from_sh(posed_means, cano_features, cano_means, camera, sh_degree, tf):
    """gaussian_utils.py:431-449 for callers that hold the materialised per-Gaussian ``tf`` [N,4,4]: the view direction is
    taken in canonical space through inv(tf).  Plain torch operations on the tensors' device, like the reference (the fused
    ``pose_gaussians`` kernel does the same arithmetic without materialising tf and is the fast path)."""
    shs_view = cano_features.transpose(1, 2).reshape(-1, 3, cano_features.shape[1])[..., : (sh_degree + 1) ** 2]
    cc = torch.as_tensor(camera.camera_center).to(posed_means.device).reshape(-1, 3)[:1].repeat(cano_features.shape[0], 1)
    if tf is not None:
        hom = torch.cat([cc, torch.ones_like(cc[:, :1])], dim=1)
        cam_inv = torch.einsum("nij,nj->ni", torch.linalg.inv(tf), hom)[..., :3]
        d = cano_means - cam_inv
    else:
        d = posed_means - cc
    d = d / d.norm(dim=1, keepdim=True)
    return torch.clamp_min(eval_sh(sh_degree, shs_view, d) + 0.5, 0.0)


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
    """gaussian_utils.py:349-428.  With ``colors_precomp=None`` the colours come from the SH features like the reference
    (:401-404): for ``tf=None`` (static object: world-space view direction) the rasterizer kernels evaluate the SH themselves
    (forward and backward), otherwise ``calculate_colors_from_sh`` runs first."""
    if _cached_screenspace:
        screenspace_points = _screenspace_leaf(posed_means)          # a leaf: .grad is populated without retain_grad()
    else:
        screenspace_points = torch.zeros_like(posed_means, dtype=posed_means.dtype, requires_grad=True, device=device) + 0
        try:
            screenspace_points.retain_grad()
        except Exception:
            pass
    shs = None
    if colors_precomp is None:
        if cano_features is None:
            raise ValueError("render_gaussians: pass cano_features (SH) or colors_precomp")
        if tf is None:
            shs = cano_features[:, : (sh_degree + 1) ** 2]         # in-kernel SH -> RGB along (mean - camera centre)
        else:
            colors_precomp = calculate_colors_from_sh(posed_means, cano_features, cano_means, camera, sh_degree, tf)
    rasterizer = GaussianRasterizer(raster_settings=_settings(camera, bg_color, sh_degree, device))
    rendered_image, radii = rasterizer(means3D=posed_means, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp,
                                       opacities=cano_opacity, scales=None, rotations=None, cov3D_precomp=posed_cov)
    rendered_image = torch.permute(rendered_image, (1, 2, 0))
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii}


def render_fused(params, skin_wts, bone_tf, camera, bg_color, sh_degree=3, isotropic=False, num_skinned=None, grad_sink=None,
                 accumulate=False):
    """params: (xyz, log_scale, quat, opacity_logit, f_dc, f_rest) -- the six nn.Parameters of GaussianModel.
    Returns the render_gaussians dict plus the posed quantities (what TrainingModule.forward returns)."""
    xyz, log_scale, quat, opacity_logit, f_dc, f_rest = params
    device = xyz.device
    campos = torch.as_tensor(camera.camera_center).to(device)
    posed_xyz, posed_cov, colors, opacity = pose_gaussians(xyz, log_scale, quat, opacity_logit, f_dc, f_rest, skin_wts, bone_tf,
                                                          campos, sh_degree, isotropic, num_skinned, grad_sink, accumulate)
    out = render_gaussians(posed_xyz, posed_cov, xyz, None, opacity, camera, bg_color, colors_precomp=colors,
                           sh_degree=sh_degree, device=device, _cached_screenspace=True)
    out.update(posed_xyz=posed_xyz, posed_cov=posed_cov, colors=colors, cano_opacity=opacity)
    return out
