"""Host camera mirror against golden outputs of the reference's get_opengl_camera_attributes."""
import os

import numpy as np

from helpers import GOLDEN
from manus_b200.cameras import opengl_camera


def test_camera_matrices_match_reference():
    g = np.load(os.path.join(GOLDEN, "camera_golden.npz"))
    for j in range(5):
        fx, fy, W, H = g[f"in_{j}"]
        cam = opengl_camera(fx, fy, g[f"extr_{j}"], int(W), int(H), dtype=np.float64)
        np.testing.assert_allclose(cam.world_view_transform, g[f"world_view_transform_{j}"], rtol=0, atol=0)
        np.testing.assert_allclose(cam.projection_matrix, g[f"projection_matrix_{j}"], rtol=1e-15, atol=1e-15)
        np.testing.assert_allclose(cam.full_proj_transform, g[f"full_proj_transform_{j}"], rtol=1e-14, atol=1e-14)
        np.testing.assert_allclose(cam.camera_center, g[f"camera_center_{j}"], rtol=1e-14, atol=1e-14)
        np.testing.assert_allclose([cam.fovx, cam.fovy], g[f"fov_{j}"], rtol=1e-15)
