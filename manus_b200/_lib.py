"""ctypes binding of the C ABI declared in include/manus_b200.h.

There is no CPU fallback: if the shared library is missing or a call fails, the op raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MANUS_B200_LIB") or os.path.join(_HERE, "lib", "libmanus_b200.so")   # env override: kernel experiments
if os.environ.get("MANUS_B200_LIB"):
    import sys as _sys
    print(f"manus_b200: MANUS_B200_LIB is set -- loading {LIB_PATH} instead of the in-tree library", file=_sys.stderr)

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
i64p = C.POINTER(C.c_int64)


class RasterInputs(C.Structure):
    """struct mb_raster_inputs"""
    _fields_ = [
        ("num_points", C.c_int32), ("image_width", C.c_int32), ("image_height", C.c_int32), ("sh_degree", C.c_int32),
        ("sh_coeffs", C.c_int32), ("prefiltered", C.c_int32), ("debug", C.c_int32),
        ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
        ("background", C.c_void_p), ("means3D", C.c_void_p), ("opacities", C.c_void_p), ("colors_precomp", C.c_void_p),
        ("shs", C.c_void_p), ("cov3D_precomp", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p),
        ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
        ("tanfov_dev", C.c_void_p),
    ]


class PoseInputs(C.Structure):
    """struct mb_pose_inputs"""
    _fields_ = [
        ("num_points", C.c_int32), ("num_skinned", C.c_int32), ("num_bones", C.c_int32), ("sh_degree", C.c_int32),
        ("sh_coeffs", C.c_int32), ("isotropic", C.c_int32),
        ("xyz", C.c_void_p), ("log_scale", C.c_void_p), ("quat", C.c_void_p), ("opacity_logit", C.c_void_p),
        ("f_dc", C.c_void_p), ("f_rest", C.c_void_p), ("skin_wts", C.c_void_p), ("bone_tf", C.c_void_p), ("campos", C.c_void_p),
        ("bones_posed", C.c_void_p), ("bones_rest_inv", C.c_void_p), ("num_posed_bones", C.c_int32), ("reserved_", C.c_int32),
    ]


class ViewInputs(C.Structure):
    """struct mb_view_inputs"""
    _fields_ = [("raster", C.c_void_p), ("radii", C.c_void_p), ("grad_scratch", C.c_void_p), ("dL_dmeans2D", C.c_void_p),
                ("bone_tf", C.c_void_p), ("bones_posed", C.c_void_p), ("campos", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/manus_b200.h declares
SIGNATURES = {
    "mb_version": (C.c_int, []),
    "mb_last_error": (C.c_char_p, []),
    "mb_device_sm_count": (C.c_int, []),
    "mb_profile_enable": (None, [C.c_int]),
    "mb_profile_report": (C.c_int, [C.c_char_p, C.c_size_t]),
    "mb_profile_timeline": (C.c_int, [C.c_char_p, C.c_size_t]),
    "mb_sh_colors_forward": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 3 + [C.c_void_p, C.c_void_p]),
    "mb_sh_colors_backward": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 3 + [C.c_void_p] * 4 + [C.c_void_p]),
    "mb_p2p_allreduce": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_void_p]),
    "mb_multimem_allreduce": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p]),
    "mb_raster_geom_bytes": (C.c_size_t, [C.c_int32]),
    "mb_raster_binning_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "mb_raster_image_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "mb_raster_forward_geom": (C.c_int, [C.POINTER(RasterInputs), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_raster_forward_render": (C.c_int, [C.POINTER(RasterInputs), C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64, C.c_void_p,
                                           C.c_size_t, C.c_void_p, C.c_void_p]),
    "mb_raster_query": (C.c_int, [C.c_void_p, i64p, i64p, i32p, C.c_void_p]),
    "mb_raster_backward_scratch_bytes": (C.c_size_t, [C.c_int32]),
    "mb_raster_backward": (C.c_int, [C.POINTER(RasterInputs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                     C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_size_t] + [C.c_void_p] * 8 + [C.c_void_p]),
    "mb_raster_state_layout": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, C.c_int32, i64p, C.c_int32]),
    "mb_mark_visible": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_pose_forward": (C.c_int, [C.POINTER(PoseInputs)] + [C.c_void_p] * 5 + [C.c_void_p]),
    "mb_pose_project_forward": (C.c_int, [C.POINTER(PoseInputs), C.POINTER(RasterInputs), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p] +
                                [C.c_void_p] * 4 + [C.c_void_p]),
    "mb_pose_backward_from_raster_views": (C.c_int, [C.POINTER(PoseInputs), C.c_int32, C.POINTER(ViewInputs)] + [C.c_void_p] * 6 +
                                           [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_pose_backward": (C.c_int, [C.POINTER(PoseInputs)] + [C.c_void_p] * 4 + [C.c_void_p] * 7 + [C.c_void_p]),
    "mb_pose_backward_from_raster": (C.c_int, [C.POINTER(PoseInputs), C.POINTER(RasterInputs), C.c_void_p, C.c_void_p, C.c_void_p] +
                                     [C.c_void_p] * 7 + [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_raster_backward_blend": (C.c_int, [C.POINTER(RasterInputs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                           C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mb_pose_backward_accumulate": (C.c_int, [C.POINTER(PoseInputs)] + [C.c_void_p] * 4 + [C.c_void_p] * 7 + [C.c_void_p]),
    "mb_skin_weights_forward": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "mb_skin_weights_backward": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_fused_adam": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_int64),
                                C.POINTER(C.c_double), C.c_int64, C.c_double, C.c_double, C.c_double, C.c_float, C.c_void_p]),
    "mb_photometric_loss_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "mb_photometric_loss": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mb_sh_grad_from_views": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_knn_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "mb_dist2_knn3": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mb_nearest_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "mb_nearest_point": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mb_sort_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "mb_radix_sort_pairs": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_size_t,
                                                        C.c_void_p]),
}

_lib = None


class ManusB200Error(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ManusB200Error(
                f"{LIB_PATH} is missing: build it with `python -m manus_b200.build` (needs nvcc). "
                "manus_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mb_last_error()
        raise ManusB200Error(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def profile_enable(on: bool) -> None:
    lib().mb_profile_enable(int(on))


def profile_timeline() -> list:
    """[(kernel, stream, start_us, duration_us)] of the recorded launches (synchronises the device; the record is kept)."""
    buf = C.create_string_buffer(1 << 20)
    check(lib().mb_profile_timeline(buf, len(buf)), "mb_profile_timeline")
    out = []
    for line in buf.value.decode().splitlines():
        name, sid, st, du = line.rsplit(" ", 3)
        out.append((name, int(sid), float(st), float(du)))
    return out


def profile_report() -> dict:
    """{kernel: (launches, total_ms)} since the last report (synchronises the device)."""
    buf = C.create_string_buffer(1 << 16)
    check(lib().mb_profile_report(buf, len(buf)), "mb_profile_report")
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.rsplit(" ", 2)
        out[name] = (int(n), float(ms))
    return out
