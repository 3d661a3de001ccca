"""The pose oracle against golden vectors computed by the reference's own Python (tests/golden/make_golden_pose.py)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN
from oracle import pose_ref as P

CASES = ["hand_voxel", "hand_points_iso", "object", "deg0", "deg1", "deg2"]


def load(name):
    g = np.load(os.path.join(GOLDEN, f"pose_golden_{name}.npz"))
    return {k: g[k] for k in g.files}


def run_oracle(g, dtype=torch.float32):
    t = lambda k: torch.tensor(g[k]).to(dtype)
    leaves = {k: t(k).requires_grad_(True) for k in ["xyz", "log_scale", "quat", "opacity_logit", "f_dc", "f_rest"]}
    B = int(g["n_bones"])
    tfs = sw = None
    if B:
        tfs = P.bone_transforms(t("bones_posed"), t("bones_rest"), B == 21)
        sw = t("skin_wts").requires_grad_(True)
        leaves["skin_wts"] = sw
    out = P.pose_gaussians_ref(leaves["xyz"], leaves["log_scale"], leaves["quat"], leaves["opacity_logit"], leaves["f_dc"],
                               leaves["f_rest"], sw, tfs, t("campos"), int(g["sh_degree"]), bool(g["isotropic"]), return_tf=True)
    return leaves, out


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference(name):
    g = load(name)
    _, out = run_oracle(g)
    for got, key in zip(out[:4], ["posed_xyz", "posed_cov", "colors", "opacity"]):
        np.testing.assert_array_equal(got.detach().numpy(), g[key], err_msg=key)   # same ops, same order: bit exact
    if int(g["n_bones"]):
        np.testing.assert_array_equal(out[4].detach().numpy(), g["tf"])


@pytest.mark.parametrize("name", CASES)
def test_backward_matches_reference(name):
    g = load(name)
    leaves, out = run_oracle(g)
    t = lambda k: torch.tensor(g[k])
    loss = (out[0] * t("G_xyz")).sum() + (out[1] * t("G_cov")).sum() + (out[2] * t("G_col")).sum() + (out[3] * t("G_op")).sum()
    names = list(leaves)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    for k, gr in zip(names, grads):
        ref = g["g_" + k]
        if ref.size == 0:
            continue
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(gr.numpy() - ref).max() <= 1e-6 * scale, k


def test_identity_pose_equals_object_path():
    """KAT: skinning with identity bones must reproduce the object path."""
    g = load("object")
    t = lambda k: torch.tensor(g[k])
    N = g["xyz"].shape[0]
    args = [t(k) for k in ["xyz", "log_scale", "quat", "opacity_logit", "f_dc", "f_rest"]]
    obj = P.pose_gaussians_ref(*args, None, None, t("campos"))
    w = torch.full((N, 3), 1.0 / 3)
    hand = P.pose_gaussians_ref(*args, w, torch.eye(4)[None].repeat(3, 1, 1), t("campos"))
    for a, b in zip(obj, hand):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)


def test_rigid_single_bone_rotates_mean_and_cov():
    """KAT: one bone, rigid transform -> x' = R x + t, Sigma' = R Sigma R^T; colour is unchanged when the camera moves along."""
    g = load("object")
    t = lambda k: torch.tensor(g[k]).double()
    N = g["xyz"].shape[0]
    ang = 0.7
    R = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1.0]], dtype=torch.float64)
    tr = torch.tensor([0.1, -0.2, 0.05], dtype=torch.float64)
    T = torch.eye(4, dtype=torch.float64); T[:3, :3] = R; T[:3, 3] = tr
    args = [t(k) for k in ["xyz", "log_scale", "quat", "opacity_logit", "f_dc", "f_rest"]]
    cam = t("campos")
    base = P.pose_gaussians_ref(*args, None, None, cam)
    moved = P.pose_gaussians_ref(*args, torch.ones(N, 1, dtype=torch.float64), T[None], R @ cam + tr)
    assert torch.allclose(moved[0], base[0] @ R.T + tr, atol=1e-12)
    S = P.build_symmetric(base[1]); S2 = P.build_symmetric(moved[1])
    assert torch.allclose(S2, R @ S @ R.T, atol=1e-14)
    assert torch.allclose(moved[2], base[2], atol=1e-9)
