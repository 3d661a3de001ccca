"""ctypes front-end for oracle/raster_ref.c (TEST INFRASTRUCTURE -- see oracle/__init__.py).

``RasterRef(precision).forward(...)`` / ``.backward(...)`` take and return numpy float32 arrays laid out
exactly like the tensors MANUS passes to ``GaussianRasterizer`` (src/utils/gaussian_utils.py:407-416).
PARITY UNPINNED (no upstream build / golden vectors exist in /root/reference); see the C file header.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")


def build(force: bool = False) -> None:
    """Compile the C restatement (gcc only; a few seconds)."""
    outs = [os.path.join(_BUILD, f"libraster_ref_{p}.so") for p in ("f32", "f64")]
    src = os.path.join(_HERE, "raster_ref.c")
    if not force and all(os.path.exists(o) and os.path.getmtime(o) >= os.path.getmtime(src) for o in outs):
        return
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class RasterRef:
    def __init__(self, precision: str = "f32"):
        assert precision in ("f32", "f64")
        build()
        self.lib = C.CDLL(os.path.join(_BUILD, f"libraster_ref_{precision}.so"))
        self.lib.raster_ref_forward.restype = C.c_void_p
        self.lib.raster_ref_free.argtypes = [C.c_void_p]
        self.ctx = None
        self.meta = None

    def close(self):
        if self.ctx:
            self.lib.raster_ref_free(self.ctx)
            self.ctx = None

    __del__ = close

    def forward(self, means3D, opacities, viewmatrix, projmatrix, campos, tanfovx, tanfovy, W, H, bg,
                colors_precomp=None, cov3D_precomp=None, shs=None, sh_degree=0, scales=None, rotations=None,
                scale_modifier=1.0, nthreads=1):
        """-> (color[3,H,W] f32, radii[N] i32, num_rendered)."""
        self.close()
        means3D = _f(means3D); N = means3D.shape[0]
        assert (shs is None) != (colors_precomp is None)
        assert (cov3D_precomp is None) != (scales is None or rotations is None)
        shs = _f(shs); M = 0 if shs is None else shs.shape[1]
        colors_precomp, cov3D_precomp, scales, rotations = _f(colors_precomp), _f(cov3D_precomp), _f(scales), _f(rotations)
        opacities = _f(opacities).reshape(-1)
        view, proj, cam, bg = _f(viewmatrix).reshape(-1), _f(projmatrix).reshape(-1), _f(campos).reshape(-1), _f(bg).reshape(-1)
        out = np.zeros((3, H, W), np.float32)
        radii = np.zeros(N, np.int32)
        D = C.c_int64(0)
        self.ctx = self.lib.raster_ref_forward(
            C.c_int(N), C.c_int(W), C.c_int(H), _p(means3D), _p(cov3D_precomp), _p(scales), _p(rotations),
            C.c_float(scale_modifier), _p(colors_precomp), _p(shs), C.c_int(sh_degree), C.c_int(M), _p(opacities),
            _p(view), _p(proj), _p(cam), C.c_float(tanfovx), C.c_float(tanfovy), _p(bg), _p(out), _p(radii, C.c_int32),
            C.byref(D), C.c_int(nthreads))
        self.meta = dict(N=N, W=W, H=H, M=M, D=D.value, shs=shs, scales=scales, rotations=rotations)
        return out, radii, D.value

    def state(self):
        m = self.meta; N, W, H = m["N"], m["W"], m["H"]
        gx, gy = (W + 15) // 16, (H + 15) // 16
        st = dict(xy=np.zeros((N, 2), np.float32), depth=np.zeros(N, np.float32), conic=np.zeros((N, 3), np.float32),
                  rgb=np.zeros((N, 3), np.float32), tiles=np.zeros(N, np.int32), final_T=np.zeros((H, W), np.float32),
                  n_contrib=np.zeros((H, W), np.int32), point_list=np.zeros(max(m["D"], 1), np.int32),
                  ranges=np.zeros((gx * gy, 2), np.int64))
        self.lib.raster_ref_get_state(C.c_void_p(self.ctx), _p(st["xy"]), _p(st["depth"]), _p(st["conic"]), _p(st["rgb"]),
                                      _p(st["tiles"], C.c_int32), _p(st["final_T"]), _p(st["n_contrib"], C.c_int32),
                                      _p(st["point_list"], C.c_int32), _p(st["ranges"], C.c_int64))
        st["point_list"] = st["point_list"][: m["D"]]
        return st

    def set_tile_sampling(self, stride=1, offset=0):
        """CPU-baseline timing only: blend (fwd and bwd) just the tiles with tile % stride == offset."""
        self.lib.raster_ref_set_tile_sampling(C.c_int(stride), C.c_int(offset))

    def backward(self, dL_dout, nthreads=1):
        """dL_dout: [3,H,W] -> dict of gradients shaped like upstream's return values.
        nthreads=1 is the deterministic checker; >1 only for CPU-baseline timing."""
        m = self.meta; N, M = m["N"], m["M"]
        g = _f(dL_dout)
        assert g.shape == (3, m["H"], m["W"])
        o = dict(means2D=np.zeros((N, 3), np.float32), colors=np.zeros((N, 3), np.float32), opacity=np.zeros((N, 1), np.float32),
                 means3D=np.zeros((N, 3), np.float32), cov3D=np.zeros((N, 6), np.float32), sh=np.zeros((N, max(M, 0), 3), np.float32),
                 scales=np.zeros((N, 3), np.float32), rotations=np.zeros((N, 4), np.float32), conic=np.zeros((N, 3), np.float32))
        self.lib.raster_ref_backward(C.c_void_p(self.ctx), _p(g), _p(m["shs"]), _p(m["scales"]), _p(m["rotations"]),
                                     _p(o["means2D"]), _p(o["colors"]), _p(o["opacity"]), _p(o["means3D"]), _p(o["cov3D"]),
                                     _p(o["sh"]) if M else None, _p(o["scales"]), _p(o["rotations"]), _p(o["conic"]),
                                     C.c_int(nthreads))
        return o

    def mark_visible(self, means3D, viewmatrix, projmatrix):
        means3D = _f(means3D); out = np.zeros(means3D.shape[0], np.uint8)
        self.lib.raster_ref_mark_visible(C.c_int(means3D.shape[0]), _p(means3D), _p(_f(viewmatrix).reshape(-1)),
                                         _p(_f(projmatrix).reshape(-1)), _p(out, C.c_uint8))
        return out.astype(bool)
