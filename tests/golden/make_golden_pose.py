"""Generate golden vectors for the pre-raster "pose" step FROM THE REFERENCE'S OWN PYTHON.

Run once in the authoring container (needs /root/reference):
    python tests/golden/make_golden_pose.py
Writes tests/golden/pose_golden_{hand_voxel,hand_points_iso,object,deg0,deg1,deg2}.npz,
tests/golden/camera_golden.npz and manus_b200/data/scene_fixtures.npz.

What runs reference code (no arithmetic is restated here):
  * src/modules/hand_dynamic.py  TrainingModule.forward   (LBS of means + covariances)
  * src/modules/object.py        TrainingModule.forward   (static object)
  * src/models/gaussian.py       GaussianModel.get_covariance / get_features / get_opacity
  * src/utils/gaussian_utils.py  calculate_colors_from_sh, build_* / strip_*
  * src/utils/sh_utils.py        eval_sh
  * src/utils/cam_utils.py       get_opengl_camera_attributes
and torch autograd through them for the gradient vectors.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import as R  # noqa: E402

R.install()
EasyDict = R.easydict_passthrough()

import joblib  # noqa: E402
import torch  # noqa: E402

import src.models.gaussian as ref_gm  # noqa: E402
import src.modules.hand_dynamic as ref_hd  # noqa: E402
import src.modules.object as ref_ob  # noqa: E402
import src.utils.cam_utils as ref_cu  # noqa: E402
import src.utils.gaussian_utils as ref_gu  # noqa: E402

torch.set_num_threads(4)
DATA = "/root/reference/data"


def make_model(N, isotropic, seed, hand_pts=None, sh_degree=3):
    """A GaussianModel carrying explicit parameters (bypasses __init__, which needs distCUDA2 on a GPU)."""
    g = torch.Generator().manual_seed(seed)
    m = object.__new__(ref_gm.GaussianModel)
    torch.nn.Module.__init__(m)
    m.opts = types.SimpleNamespace(isotropic_scaling=isotropic, sh_degree=3)
    m.setup_functions()
    if hand_pts is None:
        xyz = torch.randn(N, 3, generator=g) * 0.05
    else:
        xyz = hand_pts[torch.randperm(hand_pts.shape[0], generator=g)[:N]].float()
    m._xyz = torch.nn.Parameter(xyz.clone())
    m._scaling = torch.nn.Parameter(torch.log(torch.rand(N, 1 if isotropic else 3, generator=g) * 0.004 + 0.0005))
    m._rotation = torch.nn.Parameter(torch.randn(N, 4, generator=g) * 0.7 + torch.tensor([0.5, 0, 0, 0]))
    m._opacity = torch.nn.Parameter(torch.randn(N, 1, generator=g) * 2.0)
    m._features_dc = torch.nn.Parameter((torch.rand(N, 1, 3, generator=g) - 0.5) / 0.28209479177387814)
    m._features_rest = torch.nn.Parameter(torch.randn(N, (sh_degree + 1) ** 2 - 1, 3, generator=g) * 0.15)
    return m


def sparse_weights(N, B, seed):
    g = torch.Generator().manual_seed(seed)
    w = torch.zeros(N, B)
    for k in range(3):
        idx = torch.randint(0, B, (N,), generator=g)
        w[torch.arange(N), idx] += torch.rand(N, generator=g) + 0.05
    return w / w.sum(-1, keepdim=True)


def run_case(name, mode, N, isotropic, sh_degree, seed):
    poses = joblib.load(f"{DATA}/meta_data/novel_pose.pkl")
    cams = joblib.load(f"{DATA}/camera_paths/real.pkl")
    rest = torch.tensor(poses["rest_matrixs"]).float()
    posed = torch.tensor(poses["pose_matrixs"][100]).float()
    heads = torch.tensor(poses["rest_heads"]).float()
    tails = torch.tensor(poses["rest_tails"]).float()
    g = torch.Generator().manual_seed(seed + 1000)
    t = torch.rand(N * 2, 1, generator=g)
    b = torch.randint(0, 20, (N * 2,), generator=g)
    hand_pts = heads[b] * (1 - t) + tails[b] * t + torch.randn(N * 2, 3, generator=g) * 0.006

    K = np.array([[cams["intrs"][7][0], 0, cams["intrs"][7][2]], [0, cams["intrs"][7][1], cams["intrs"][7][3]], [0, 0, 1.0]])
    cam = ref_cu.get_opengl_camera_attributes(K, np.array(cams["extrs"][7]), 1920, 1080)
    campos = torch.tensor(cam["camera_center"]).float()
    camera = types.SimpleNamespace(camera_center=campos)

    model = make_model(N, isotropic, seed, hand_pts if mode != "object" else None, sh_degree)
    out = {}
    if mode == "object":
        fake = types.SimpleNamespace(model=model)
        pred = ref_ob.TrainingModule.forward(fake, {})
        skin = None
        B = 0
    else:
        voxel = mode == "hand_voxel"
        B = 21 if voxel else 20
        skin = sparse_weights(N, B, seed + 5).requires_grad_(True)
        model.get_skin_weights = lambda: skin
        opts = types.SimpleNamespace(model=types.SimpleNamespace(opts=types.SimpleNamespace(
            skin_weights_init_type="mano_init_voxel" if voxel else "mano_init_points")))
        fake = types.SimpleNamespace(model=model, opts=opts)
        batch = {"bones_posed": types.SimpleNamespace(transforms=posed),
                 "bones_rest": types.SimpleNamespace(transforms=rest)}
        pred = ref_hd.TrainingModule.forward(fake, batch)
        out["bones_posed"] = posed.numpy()
        out["bones_rest"] = rest.numpy()
        out["skin_wts"] = skin.detach().numpy()
        out["tf"] = pred["tf"].detach().numpy()

    colors = ref_gu.calculate_colors_from_sh(pred["posed_xyz"], pred["cano_features"], pred["cano_xyz"],
                                             camera, sh_degree, pred["tf"])
    gg = torch.Generator().manual_seed(seed + 77)
    G_xyz = torch.rand(N, 3, generator=gg) - 0.5
    G_cov = (torch.rand(N, 6, generator=gg) - 0.5) * 1e3
    G_col = torch.rand(N, 3, generator=gg) - 0.5
    G_op = torch.rand(N, 1, generator=gg) - 0.5
    loss = ((pred["posed_xyz"] * G_xyz).sum() + (pred["posed_cov"] * G_cov).sum()
            + (colors * G_col).sum() + (pred["cano_opacity"] * G_op).sum())
    params = [model._xyz, model._scaling, model._rotation, model._opacity, model._features_dc, model._features_rest]
    if skin is not None:
        params.append(skin)
    grads = torch.autograd.grad(loss, params, allow_unused=True)

    out.update(
        xyz=model._xyz.detach().numpy(), log_scale=model._scaling.detach().numpy(),
        quat=model._rotation.detach().numpy(), opacity_logit=model._opacity.detach().numpy(),
        f_dc=model._features_dc.detach().numpy(), f_rest=model._features_rest.detach().numpy(),
        campos=campos.numpy(), sh_degree=np.int32(sh_degree), isotropic=np.int32(isotropic), n_bones=np.int32(B),
        posed_xyz=pred["posed_xyz"].detach().numpy(), posed_cov=pred["posed_cov"].detach().numpy(),
        colors=colors.detach().numpy(), opacity=pred["cano_opacity"].detach().numpy(),
        G_xyz=G_xyz.numpy(), G_cov=G_cov.numpy(), G_col=G_col.numpy(), G_op=G_op.numpy(),
        g_xyz=grads[0].numpy(), g_log_scale=grads[1].numpy(), g_quat=grads[2].numpy(),
        g_opacity_logit=grads[3].numpy(), g_f_dc=grads[4].numpy(), g_f_rest=grads[5].numpy(),
    )
    if skin is not None:
        out["g_skin_wts"] = grads[6].numpy()
    out = {k: (v.astype(np.float32) if isinstance(v, np.ndarray) and v.dtype == np.float64 else v) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, f"pose_golden_{name}.npz"), **out)
    print(name, "N", N, "B", B, "colors mean", float(colors.detach().mean()), "clamped frac", float((colors.detach() == 0).float().mean()))


def camera_golden():
    """get_opengl_camera_attributes (src/utils/cam_utils.py:50-78) on five shipped cameras at two resolutions."""
    cams = joblib.load(f"{DATA}/camera_paths/real.pkl")
    out = {}
    for j, (ci, W, H) in enumerate([(0, 1920, 1080), (7, 1920, 1080), (100, 1920, 1080), (200, 1920, 1080), (0, 800, 800)]):
        fx, fy, cx, cy = cams["intrs"][ci]
        if (W, H) != (1920, 1080):
            fx, fy = fx * W / 1920.0, fy * H / 1080.0
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
        a = ref_cu.get_opengl_camera_attributes(K, np.array(cams["extrs"][ci]), W, H)
        out[f"in_{j}"] = np.array([fx, fy, W, H], dtype=np.float64)
        out[f"extr_{j}"] = np.array(cams["extrs"][ci], dtype=np.float64)
        for k in ("world_view_transform", "projection_matrix", "full_proj_transform", "camera_center"):
            out[f"{k}_{j}"] = np.asarray(a[k], dtype=np.float64)
        out[f"fov_{j}"] = np.array([a["fovx"], a["fovy"]], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "camera_golden.npz"), **out)


def scene_fixtures():
    """Small derived copy of the shipped data pickles so that the bench and the
    GPU tests can build realistic scenes without /root/reference: 51 cameras/poses (every 5th), rest bones,
    MANO rest vertices + 16-joint weights."""
    cams = joblib.load(f"{DATA}/camera_paths/real.pkl")
    poses = joblib.load(f"{DATA}/meta_data/novel_pose.pkl")
    mano = joblib.load(f"{DATA}/mano/mano_rest.pkl")
    sel = np.arange(0, 251, 5)
    np.savez_compressed(
        os.path.join(HERE, "..", "..", "manus_b200", "data", "scene_fixtures.npz"),
        cam_intrs=np.asarray(cams["intrs"])[sel].astype(np.float64),
        cam_extrs=np.asarray(cams["extrs"])[sel].astype(np.float64),
        rest_matrixs=np.asarray(poses["rest_matrixs"]).astype(np.float64),
        rest_heads=np.asarray(poses["rest_heads"]).astype(np.float64),
        rest_tails=np.asarray(poses["rest_tails"]).astype(np.float64),
        pose_matrixs=np.asarray(poses["pose_matrixs"])[sel].astype(np.float64),
        mano_vert=np.asarray(mano["vert"]).astype(np.float32),
        mano_weights=np.asarray(mano["weights"]).astype(np.float32),
        frame_index=sel.astype(np.int32),
    )


if __name__ == "__main__":
    run_case("hand_voxel", "hand_voxel", 384, False, 3, 11)
    run_case("hand_points_iso", "hand_points", 256, True, 3, 12)
    run_case("object", "object", 384, False, 3, 13)
    run_case("deg0", "hand_voxel", 64, False, 0, 14)
    run_case("deg1", "object", 64, False, 1, 15)
    run_case("deg2", "hand_voxel", 64, False, 2, 16)
    camera_golden()
    scene_fixtures()
    print("done")
