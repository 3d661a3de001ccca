"""Seeded synthetic scenes that mirror MANUS's own initialisation (bench / test harness; numpy only).

Recipe = SURVEY.md Appendix C:
  * hand Gaussians sampled on the rest bones exactly like ``sample_gaussians_on_bones_func``
    (/root/reference/src/utils/train_utils.py:104-139): S per bone from N(mid, R diag(len/5,len/4,len/4)^2 R^T) and
    S/2 per joint from N(head, R diag(len/6,len/4,len/6)^2 R^T)  -> 20*1.5*S Gaussians (S=10000 -> 300k,
    scripts/train/train_hands.sh:31);
  * skin weights like ``init_mano_weights`` (train_utils.py:48-84): MANO 16-joint weights remapped to the 20 bones
    (``mano_to_ours``, :68), averaged over the 20 nearest MANO rest vertices (:71-74), plus the identity "background"
    column of the voxel mode (src/modules/hand_dynamic.py:98-102);
  * colours U[0,1] -> RGB2SH in f_dc (src/models/gaussian.py:103-106), small f_rest noise so degree-3 terms count;
  * log-scales from the mean squared distance to the 3 nearest neighbours (gaussian.py:110-114);
  * quaternions (1,0,0,0) (+ noise), opacity logit of 0.1 (gaussian.py:118) perturbed so early termination occurs;
  * object: points near a 0.15 m blob placed at the palm.
Bones, poses, cameras and MANO data come from ``manus_b200/data/scene_fixtures.npz`` (a small derived copy of the
pickles shipped in /root/reference/data, written by tests/golden/make_golden_pose.py).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from .cameras import Camera, opengl_camera

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "scene_fixtures.npz")
C0 = 0.28209479177387814
MANO_TO_OURS = [13, 14, 14, 15, 0, 1, 2, 3, 0, 4, 5, 6, 0, 10, 11, 12, 0, 7, 8, 9]  # train_utils.py:68


def fixtures():
    return np.load(_DATA)


def knn3_mean_sq(points: np.ndarray) -> np.ndarray:
    """Harness-side exact 3-NN mean squared distance (scene construction only; the product op is distCUDA2)."""
    from scipy.spatial import cKDTree

    d, _ = cKDTree(points.astype(np.float64)).query(points.astype(np.float64), k=4)
    return (d[:, 1:] ** 2).mean(1).astype(np.float32)


@dataclass
class Scene:
    xyz: np.ndarray            # [N,3]
    log_scale: np.ndarray      # [N,3]
    quat: np.ndarray           # [N,4]
    opacity_logit: np.ndarray  # [N,1]
    f_dc: np.ndarray           # [N,1,3]
    f_rest: np.ndarray         # [N,15,3]
    n_hand: int                # the first n_hand Gaussians are skinned, the rest are static (tf = I)
    skin_wts: np.ndarray | None  # [n_hand,21]
    bones_rest: np.ndarray = field(default=None)     # [20,4,4]
    meta: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return self.xyz.shape[0]


def _sample_hand_points(rng, S: int, fx) -> np.ndarray:
    heads, tails, R = fx["rest_heads"], fx["rest_tails"], fx["rest_matrixs"][:, :3, :3]
    length = np.linalg.norm(tails - heads, axis=1, keepdims=True)
    pts = []
    for centre, div, cnt in (((heads + tails) / 2, (5.0, 4.0, 4.0), S), (heads, (6.0, 4.0, 6.0), S // 2)):
        scale = np.concatenate([length / div[0], length / div[1], length / div[2]], axis=1)      # [20,3]
        z = rng.standard_normal((cnt, 20, 3))
        local = z * scale[None]
        world = np.einsum("bij,nbj->nbi", R, local) + centre[None]
        pts.append(world.reshape(-1, 3))
    return np.concatenate(pts, 0).astype(np.float32)


def hand_skin_weights(points: np.ndarray, fx, neighbors: int = 20, background: bool = True) -> np.ndarray:
    from scipy.spatial import cKDTree

    w20 = fx["mano_weights"][:, MANO_TO_OURS]
    _, idx = cKDTree(fx["mano_vert"].astype(np.float64)).query(points.astype(np.float64), k=neighbors)
    w = w20[idx].mean(1)
    if background:
        w = np.concatenate([w, np.zeros((w.shape[0], 1), w.dtype)], -1)
    w = w / w.sum(-1, keepdims=True)
    return w.astype(np.float32)


def _appearance(rng, n: int, points: np.ndarray, quat_noise: float, sh_noise: float):
    d2 = np.maximum(knn3_mean_sq(points), 1e-7)
    log_scale = np.repeat(np.log(np.sqrt(d2))[:, None], 3, 1).astype(np.float32)
    log_scale += (rng.standard_normal((n, 3)) * 0.15).astype(np.float32)          # mild anisotropy, as after training
    quat = np.zeros((n, 4), np.float32); quat[:, 0] = 1
    quat += (rng.standard_normal((n, 4)) * quat_noise).astype(np.float32)
    logit01 = np.log(0.1 / 0.9)
    opacity_logit = (logit01 + rng.uniform(-2.0, 4.0, (n, 1))).astype(np.float32)
    f_dc = ((rng.uniform(0, 1, (n, 1, 3)) - 0.5) / C0).astype(np.float32)
    f_rest = (rng.standard_normal((n, 15, 3)) * sh_noise).astype(np.float32)
    return log_scale, quat, opacity_logit, f_dc, f_rest


def make_hand(n: int, seed: int = 0, quat_noise: float = 0.1, sh_noise: float = 0.05) -> Scene:
    fx = fixtures()
    rng = np.random.default_rng(seed)
    S = int(np.ceil(n / 30.0))
    pts = _sample_hand_points(rng, S + (S % 2), fx)
    pts = pts[rng.permutation(pts.shape[0])[:n]]
    ls, q, ol, fdc, fr = _appearance(rng, n, pts, quat_noise, sh_noise)
    return Scene(pts, ls, q, ol, fdc, fr, n, hand_skin_weights(pts, fx), fx["rest_matrixs"].astype(np.float32),
                 dict(kind="hand", seed=seed))


def make_object(n: int, seed: int = 1, centre=(0.0, 0.0, 0.09), radius: float = 0.15, quat_noise: float = 0.1,
                sh_noise: float = 0.05) -> Scene:
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = radius * (0.35 + 0.65 * rng.uniform(0, 1, (n, 1)) ** (1 / 3))
    pts = (np.asarray(centre)[None] + d * r * np.array([0.55, 0.35, 0.45])[None]).astype(np.float32)
    pts += (rng.standard_normal((n, 3)) * 0.03 * radius * 0.1).astype(np.float32)
    ls, q, ol, fdc, fr = _appearance(rng, n, pts, quat_noise, sh_noise)
    return Scene(pts, ls, q, ol, fdc, fr, 0, None, None, dict(kind="object", seed=seed))


def make_composite(n: int, seed: int = 0, hand_frac: float = 0.6) -> Scene:
    """Hand Gaussians first (skinned), then object Gaussians (tf = I) -- src/modules/composite.py:50-60."""
    nh = int(round(n * hand_frac))
    h, o = make_hand(nh, seed), make_object(n - nh, seed + 1, centre=(-0.01, 0.0, 0.10), radius=0.10)
    cat = lambda a, b: np.concatenate([a, b], 0)
    return Scene(cat(h.xyz, o.xyz), cat(h.log_scale, o.log_scale), cat(h.quat, o.quat), cat(h.opacity_logit, o.opacity_logit),
                 cat(h.f_dc, o.f_dc), cat(h.f_rest, o.f_rest), nh, h.skin_wts, h.bones_rest, dict(kind="composite", seed=seed))


def camera(view: int, width: int = 1920, height: int = 1080) -> Camera:
    """One of the 51 shipped cameras (every 5th of data/camera_paths/real.pkl), rescaled like resize_factor does."""
    fx = fixtures()
    f_x, f_y, _, _ = fx["cam_intrs"][view % fx["cam_intrs"].shape[0]]
    return opengl_camera(f_x * width / 1920.0, f_y * height / 1080.0, fx["cam_extrs"][view % fx["cam_extrs"].shape[0]],
                         width, height)


def posed_bones(view: int) -> np.ndarray:
    fx = fixtures()
    return fx["pose_matrixs"][view % fx["pose_matrixs"].shape[0]].astype(np.float32)
