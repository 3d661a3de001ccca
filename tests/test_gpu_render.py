"""``render_gaussians`` called the way MANUS calls it when the colours are NOT precomputed (gaussian_utils.py:401-404):
static object (tf=None: the rasterizer kernels evaluate the SH along mean - camera centre, forward and backward) and
articulated hand (tf given: canonical view direction through inv(tf)).  Checked against the pinned pose oracle's colours
fed through the colours-precomputed path (itself checked against the rasterizer oracle in test_gpu_raster.py)."""
import numpy as np
import pytest
import torch

from helpers import GRAD_RTOL, grad_close, zoom_camera
from manus_b200 import synth
from oracle import pose_ref

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _leaf(a):
    return torch.tensor(np.ascontiguousarray(a), device=DEV).requires_grad_(True)


def _run(colors_from, sc, cam, view, G):
    """One forward + backward of render_gaussians; ``colors_from`` in {"kernel", "oracle"} -> (image, grads dict)."""
    from manus_b200.render import render_gaussians

    t = lambda a: torch.tensor(a, device=DEV)
    xyz, feats = _leaf(sc.xyz), _leaf(np.concatenate([sc.f_dc, sc.f_rest], 1))
    op = torch.sigmoid(t(sc.opacity_logit) + 1.0)
    cov6 = pose_ref.get_covariance(t(sc.log_scale) + 0.3, t(sc.quat), False, full=sc.n_hand > 0)
    tf = None
    posed, cov = xyz, cov6
    if sc.n_hand > 0:
        tfs = pose_ref.bone_transforms(t(synth.posed_bones(view)), t(sc.bones_rest), True)
        tf = torch.einsum("nb,bij->nij", t(sc.skin_wts), tfs)
        posed = torch.einsum("nij,nj->ni", tf, torch.cat([xyz, torch.ones_like(xyz[:, :1])], 1))[..., :3]
        R = tf[:, :3, :3]
        cov = pose_ref.strip_symmetric(R @ cov6 @ R.transpose(1, 2))
    colors = None
    if colors_from == "oracle":
        colors = pose_ref.calculate_colors_from_sh(posed, feats, xyz, t(cam.camera_center), 3, tf)
    out = render_gaussians(posed, cov, xyz, feats, op, cam, t(np.ones(3, np.float32)), colors_precomp=colors, sh_degree=3, tf=tf,
                           device=torch.device(DEV))
    assert out["render"].shape == (cam.height, cam.width, 3) and out["visibility_filter"].dtype == torch.bool
    (out["render"] * G).sum().backward()
    return out["render"].detach().cpu().numpy(), dict(xyz=xyz.grad.cpu().numpy(), features=feats.grad.cpu().numpy(),
                                                      means2D=out["viewspace_points"].grad.cpu().numpy())


@pytest.mark.parametrize("kind", ["object", "hand"])
def test_render_gaussians_without_precomputed_colors(built_lib, kind):
    sc = synth.make_object(3000, seed=4, centre=(0.0, 0.0, 0.05), radius=0.08) if kind == "object" else synth.make_hand(3000, seed=4)
    cam = zoom_camera(11, 160, 120, 1.6)
    G = torch.rand(120, 160, 3, generator=torch.Generator().manual_seed(9)).to(DEV)
    img_k, g_k = _run("kernel", sc, cam, 13, G)
    img_o, g_o = _run("oracle", sc, cam, 13, G)
    assert float(np.abs(img_k - 1.0).max()) > 0.1                      # something was drawn
    np.testing.assert_allclose(img_k, img_o, atol=1e-5, rtol=0)
    for k in g_k:
        ok_, e, s = grad_close(g_k[k], g_o[k], GRAD_RTOL)
        assert ok_, (kind, k, e, s)
    assert float(np.abs(g_k["features"]).max()) > 0


def test_render_gaussians_needs_colors_or_features(built_lib):
    from manus_b200.render import render_gaussians

    cam = zoom_camera(0, 64, 48)
    z = torch.zeros(4, 3, device=DEV)
    with pytest.raises(ValueError):
        render_gaussians(z, torch.zeros(4, 6, device=DEV), z, None, torch.zeros(4, 1, device=DEV), cam, torch.ones(3, device=DEV))


@pytest.mark.parametrize("with_tf", [True, False])
def test_calculate_colors_from_sh_kernel_matches_the_reference_ops(built_lib, with_tf):
    """calculate_colors_from_sh with a materialised tf (gaussian_utils.py:431-449): the kernel's closed-form 4x4 inverse and SH
    evaluation against the restated reference ops (torch.linalg.inv + eval_sh, autograd), forward and the gradients to the
    features, the means and tf itself; general invertible tf (blends of rigid transforms plus a perturbation of every entry)."""
    import types

    from manus_b200.render import calculate_colors_from_sh

    sc = synth.make_hand(2500, seed=6)
    t = lambda a: torch.tensor(a, device=DEV)
    tfs = pose_ref.bone_transforms(t(synth.posed_bones(21)), t(sc.bones_rest), True)
    tf0 = torch.einsum("nb,bij->nij", t(sc.skin_wts), tfs)
    tf0 = tf0 + 0.02 * torch.randn(tf0.shape, generator=torch.Generator().manual_seed(1)).to(DEV)
    cam = types.SimpleNamespace(camera_center=np.array([0.3, -0.2, 1.4], np.float32))
    feats0 = np.concatenate([sc.f_dc, sc.f_rest * 4.0], 1)
    G = torch.rand(sc.n, 3, generator=torch.Generator().manual_seed(2)).to(DEV)
    res = {}
    for which in ("kernel", "oracle"):
        xyz, feats = _leaf(sc.xyz), _leaf(feats0)
        tf = tf0.clone().requires_grad_(True) if with_tf else None
        posed = xyz if tf is None else torch.einsum("nij,nj->ni", tf.detach(), torch.cat([xyz.detach(), torch.ones_like(xyz[:, :1])], 1))[..., :3]
        if which == "kernel":
            col = calculate_colors_from_sh(posed, feats, xyz, cam, 3, tf)
        else:
            col = pose_ref.calculate_colors_from_sh(posed, feats, xyz, t(cam.camera_center), 3, tf)
        (col * G).sum().backward()
        res[which] = (col.detach().cpu().numpy(), xyz.grad.cpu().numpy(), feats.grad.cpu().numpy(), None if tf is None else tf.grad.cpu().numpy())
    np.testing.assert_allclose(res["kernel"][0], res["oracle"][0], atol=2e-6, rtol=0)
    assert float((res["kernel"][0] == 0).mean()) > 0.001            # the clamp is exercised
    for k in (1, 2, 3):
        if res["oracle"][k] is None:
            continue
        ok_, e, s = grad_close(res["kernel"][k], res["oracle"][k], GRAD_RTOL)
        assert ok_, (with_tf, k, e, s)
