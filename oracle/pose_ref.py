"""Oracle for the pre-raster "pose" step.

TEST INFRASTRUCTURE -- never imported by the product (see oracle/__init__.py).

A device/dtype-parametrised PyTorch restatement of what MANUS does in Python
between "parameters + bones + camera" and the rasterizer call.  Every function
cites the reference lines it follows (paths relative to /root/reference).  The
reference versions hard-code ``device="cuda"`` (src/utils/gaussian_utils.py:249,
279,305), so they cannot run on a CPU as written; the arithmetic and its order
are kept so that autograd on this file gives the reference gradients.

Pinned by tests/golden/pose_golden_*.npz (generated from the reference's own
functions by tests/golden/make_golden_pose.py).
"""
from __future__ import annotations

import torch

# src/utils/sh_utils.py:26-55
C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
      -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
      -0.4570457994644658, 1.445305721320277, -0.5900435899266435]
C4 = [2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892,
      0.10578554691520431, -0.6690465435572892, 0.47308734787878004, -1.7701307697799304,
      0.6258357354491761]


def homo(points: torch.Tensor) -> torch.Tensor:
    """src/utils/extra.py:245-246."""
    return torch.nn.functional.pad(points, (0, 1), value=1)


def bone_transforms(posed: torch.Tensor, rest: torch.Tensor, append_identity: bool) -> torch.Tensor:
    """T_b = posed_b . inv(rest_b), plus the identity "background" bone.

    src/modules/hand_dynamic.py:93-102.  posed, rest: [B,4,4] -> [B(+1),4,4].
    """
    tfs = torch.einsum("nij,njk->nik", posed, torch.linalg.inv(rest))
    if append_identity:
        tfs = torch.cat([tfs, torch.eye(4, dtype=tfs.dtype, device=tfs.device)[None]], dim=0)
    return tfs


def build_rotation(r: torch.Tensor) -> torch.Tensor:
    """src/utils/gaussian_utils.py:279-302 (quaternion normalised inside, order r,x,y,z)."""
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rows = [
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y),
    ]
    return torch.stack(rows, dim=-1).reshape(-1, 3, 3)


def build_scaling_rotation(s: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
    """L = R . diag(s); src/utils/gaussian_utils.py:305-314."""
    return build_rotation(r) @ torch.diag_embed(s)


def strip_symmetric(sym: torch.Tensor) -> torch.Tensor:
    """6-vector (xx,xy,xz,yy,yz,zz); src/utils/gaussian_utils.py:248-261."""
    return torch.stack([sym[:, 0, 0], sym[:, 0, 1], sym[:, 0, 2], sym[:, 1, 1], sym[:, 1, 2], sym[:, 2, 2]], dim=-1)


def build_symmetric(v: torch.Tensor) -> torch.Tensor:
    """src/utils/gaussian_utils.py:264-276."""
    rows = [v[:, 0], v[:, 1], v[:, 2], v[:, 1], v[:, 3], v[:, 4], v[:, 2], v[:, 4], v[:, 5]]
    return torch.stack(rows, dim=-1).reshape(-1, 3, 3)


def get_covariance(log_scale: torch.Tensor, quat: torch.Tensor, isotropic: bool, full: bool,
                   scaling_modifier: float = 1.0) -> torch.Tensor:
    """src/models/gaussian.py:48-53,84-93 (exp activation :55,64; raw _rotation passed :91,93)."""
    scaling = torch.exp(log_scale)
    if isotropic:
        scaling = scaling.repeat(1, 3)
    L = build_scaling_rotation(scaling_modifier * scaling, quat)
    cov6 = strip_symmetric(L @ L.transpose(1, 2))
    return build_symmetric(cov6) if full else cov6


def eval_sh(deg: int, sh: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """src/utils/sh_utils.py:57-120.  sh: [N,3,K], dirs: [N,3] -> [N,3]."""
    assert 0 <= deg <= 4 and sh.shape[-1] >= (deg + 1) ** 2
    result = C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        result = result - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            result = (result + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5]
                      + C2[2] * (2.0 * zz - xx - yy) * sh[..., 6] + C2[3] * xz * sh[..., 7]
                      + C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                result = (result + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10]
                          + C3[2] * y * (4 * zz - xx - yy) * sh[..., 11]
                          + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
                          + C3[4] * x * (4 * zz - xx - yy) * sh[..., 13]
                          + C3[5] * z * (xx - yy) * sh[..., 14] + C3[6] * x * (xx - 3 * yy) * sh[..., 15])
                if deg > 3:
                    result = (result + C4[0] * xy * (xx - yy) * sh[..., 16]
                              + C4[1] * yz * (3 * xx - yy) * sh[..., 17]
                              + C4[2] * xy * (7 * zz - 1) * sh[..., 18]
                              + C4[3] * yz * (7 * zz - 3) * sh[..., 19]
                              + C4[4] * (zz * (35 * zz - 30) + 3) * sh[..., 20]
                              + C4[5] * xz * (7 * zz - 3) * sh[..., 21]
                              + C4[6] * (xx - yy) * (7 * zz - 1) * sh[..., 22]
                              + C4[7] * xz * (xx - 3 * yy) * sh[..., 23]
                              + C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy)) * sh[..., 24])
    return result


def calculate_colors_from_sh(posed_means, cano_features, cano_means, camera_center, sh_degree, tf):
    """src/utils/gaussian_utils.py:431-449.  camera_center: [3] (or [1,3])."""
    shs_view = cano_features.transpose(1, 2).reshape(-1, 3, cano_features.shape[1])[..., : (sh_degree + 1) ** 2]
    cc = camera_center.reshape(-1, 3)[:1].repeat(cano_features.shape[0], 1)
    if tf is not None:
        cam_inv = torch.einsum("nij, nj->ni", torch.linalg.inv(tf), homo(cc))[..., :3]
        d = cano_means - cam_inv
    else:
        d = posed_means - cc
    d = d / d.norm(dim=1, keepdim=True)
    return torch.clamp_min(eval_sh(sh_degree, shs_view, d) + 0.5, 0.0)


def pose_gaussians_ref(xyz, log_scale, quat, opacity_logit, f_dc, f_rest, skin_wts, bone_tf, campos,
                       sh_degree: int = 3, isotropic: bool = False, return_tf: bool = False):
    """The whole pre-raster step for one frame.

    hand (skin_wts/bone_tf given): src/modules/hand_dynamic.py:86-137
    object (both None):            src/modules/object.py:32-41
    colours:                       src/utils/gaussian_utils.py:401-404,431-449
    activations:                   src/models/gaussian.py:55-82

    bone_tf is the [B,4,4] stack *after* ``bone_transforms`` (identity bone included when used).
    Returns (posed_xyz[N,3], posed_cov6[N,6], colors[N,3], opacity[N,1][, tf[N,4,4] | None]).
    """
    features = torch.cat((f_dc, f_rest), dim=1)                 # gaussian.py:73-76
    opacity = torch.sigmoid(opacity_logit)                      # gaussian.py:78-80
    if skin_wts is not None:
        assert skin_wts.shape[-1] == bone_tf.shape[0]           # hand_dynamic.py:104
        tf = torch.einsum("nb, bij->nij", skin_wts, bone_tf)    # :106
        posed_xyz = torch.einsum("nij, nj->ni", tf, homo(xyz))[..., :3]          # :107
        cov = get_covariance(log_scale, quat, isotropic, full=True)              # :123
        R = tf[..., :3, :3]
        cov = torch.einsum("bij,bjk,bkl->bil", R, cov, R.transpose(1, 2))        # :124-126
        cov6 = strip_symmetric(cov)                                              # :127
    else:
        tf = None
        posed_xyz = xyz
        cov6 = get_covariance(log_scale, quat, isotropic, full=False)            # object.py:35
    colors = calculate_colors_from_sh(posed_xyz, features, xyz, campos, sh_degree, tf)
    out = (posed_xyz, cov6, colors, opacity)
    return out + (tf,) if return_tf else out
