"""Camera attributes exactly as MANUS hands them to the rasterizer.

Host-side mirror of ``get_opengl_camera_attributes`` / ``getProjectionMatrix`` / ``focal2fov``
(/root/reference/src/utils/cam_utils.py:19-78): row-vector ("transposed") 4x4 matrices computed in float64
and converted to float32 at the boundary, principal point ignored (fov-only, centred), znear=0.01, zfar=100.
Checked against tests/golden/camera_golden.npz (generated from the reference function).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


def focal2fov(focal: float, pixels: float) -> float:
    """cam_utils.py:46-47"""
    return 2 * math.atan(pixels / (2 * focal))


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> np.ndarray:
    """cam_utils.py:19-39 (column-vector form; transposed by the caller)."""
    tan_y, tan_x = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = tan_y * znear, tan_x * znear
    P = np.zeros((4, 4))
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


@dataclass
class Camera:
    """The fields of ``Cameras`` (src/utils/structures.py) that the render path reads."""
    width: int
    height: int
    fovx: float
    fovy: float
    world_view_transform: np.ndarray   # [4,4] float32, = extr_4x4^T
    full_proj_transform: np.ndarray    # [4,4] float32, = (P . extr_4x4)^T
    camera_center: np.ndarray          # [3]   float32
    projection_matrix: np.ndarray      # [4,4] float32, = P^T
    tanfov_dev: object = None          # optional CUDA tensor [2] = (tan(fovx/2), tan(fovy/2)); used instead of fovx / fovy

    @property
    def tanfovx(self) -> float:
        return math.tan(self.fovx * 0.5)

    @property
    def tanfovy(self) -> float:
        return math.tan(self.fovy * 0.5)


def opengl_camera(fx: float, fy: float, extr_3x4, width: int, height: int, zfar: float = 100.0,
                  znear: float = 0.01, dtype=np.float32) -> Camera:
    """cam_utils.py:50-78 with resize_factor already applied to (fx, fy, width, height)."""
    fovx, fovy = focal2fov(fx, width), focal2fov(fy, height)
    extr = np.concatenate([np.asarray(extr_3x4, dtype=np.float64), np.array([[0, 0, 0, 1.0]])], axis=0)
    wv = extr.T
    pm = projection_matrix(znear, zfar, fovx, fovy).T
    full = wv @ pm
    center = np.linalg.inv(wv)[3, :3]
    return Camera(int(width), int(height), fovx, fovy, wv.astype(dtype), full.astype(dtype), center.astype(dtype),
                  pm.astype(dtype))
