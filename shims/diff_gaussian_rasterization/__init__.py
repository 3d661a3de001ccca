"""Drop-in shim: put <repo>/shims on PYTHONPATH (ahead of any upstream install) and MANUS's
``from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer``
(/root/reference/src/utils/gaussian_utils.py:18-21) resolves to the B200-native implementation."""
from manus_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                    rasterize_gaussians)
from . import _C  # noqa: F401,E402
