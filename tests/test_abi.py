"""The C-ABI library builds, loads without a GPU and exports every symbol include/manus_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "manus_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ["mb_raster_forward_geom", "mb_raster_forward_render", "mb_raster_backward", "mb_mark_visible", "mb_pose_forward",
                 "mb_pose_backward", "mb_dist2_knn3", "mb_radix_sort_pairs"]:
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    from manus_b200 import _lib

    for name in declared_symbols():
        assert hasattr(built_lib, name), f"{name} not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert built_lib.mb_version() >= 100


def test_workspace_size_queries(built_lib):
    assert built_lib.mb_raster_geom_bytes(500_000) > 500_000 * 60
    assert built_lib.mb_raster_binning_bytes(1_000_000, 1920, 1080) > 1_000_000 * 16
    assert built_lib.mb_raster_image_bytes(1920, 1080) >= 1920 * 1080 * 8
    assert built_lib.mb_knn_workspace_bytes(1000) > 0 and built_lib.mb_sort_workspace_bytes(1 << 20) > 0


def test_argument_validation_without_gpu(built_lib):
    """Validation errors are reported through return codes + mb_last_error, before any CUDA call."""
    from manus_b200 import _lib

    ri = _lib.RasterInputs()
    ri.num_points, ri.image_width, ri.image_height = 4, 0, 16
    rc = built_lib.mb_raster_forward_geom(ctypes.byref(ri), None, 0, None, None, None)
    assert rc == 1 and b"bad sizes" in built_lib.mb_last_error()
    pi = _lib.PoseInputs()
    pi.num_points, pi.num_skinned = 4, 9
    rc = built_lib.mb_pose_forward(ctypes.byref(pi), None, None, None, None, None, None)
    assert rc == 1 and b"bad counts" in built_lib.mb_last_error()


def test_product_has_no_cpu_fallback():
    """Ops must fail loudly on CPU tensors instead of computing somewhere else; the package never imports oracle/."""
    import torch

    from manus_b200 import _lib, knn, pose
    from manus_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer

    with pytest.raises(_lib.ManusB200Error):
        knn.distCUDA2(torch.zeros(8, 3))
    with pytest.raises(_lib.ManusB200Error):
        pose.pose_gaussians(torch.zeros(2, 3), torch.zeros(2, 3), torch.ones(2, 4), torch.zeros(2, 1), torch.zeros(2, 1, 3),
                            torch.zeros(2, 15, 3), None, None, torch.zeros(3))
    rs = GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 3, torch.zeros(3), False, False)
    with pytest.raises(_lib.ManusB200Error):
        GaussianRasterizer(rs)(torch.zeros(2, 3), torch.zeros(2, 3), torch.ones(2, 1), colors_precomp=torch.zeros(2, 3),
                               cov3D_precomp=torch.zeros(2, 6))
    pkg = os.path.join(ROOT, "manus_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_rasterizer_argument_errors_match_upstream():
    import torch

    from manus_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer

    rs = GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 3, torch.zeros(3), False, False)
    r = GaussianRasterizer(rs)
    z = torch.zeros(2, 3)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(z, z, torch.ones(2, 1), cov3D_precomp=torch.zeros(2, 6))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(z, z, torch.ones(2, 1), colors_precomp=z)
    assert GaussianRasterizationSettings._fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier",
                                                     "viewmatrix", "projmatrix", "sh_degree", "campos", "prefiltered", "debug")


def test_shims_expose_the_upstream_module_surface():
    """What MANUS imports (gaussian_utils.py:18-21, gaussian.py:4) plus the function-level ``_C`` surface of SURVEY.md 8b, with
    upstream's positional parameter order."""
    import inspect
    import sys

    sys.path.insert(0, os.path.join(ROOT, "shims"))
    try:
        import diff_gaussian_rasterization as dgr
        from simple_knn._C import distCUDA2  # noqa: F401
    finally:
        sys.path.pop(0)
    assert dgr.GaussianRasterizationSettings._fields[0] == "image_height" and callable(dgr.GaussianRasterizer)
    fwd = list(inspect.signature(dgr._C.rasterize_gaussians).parameters)
    assert fwd == ["background", "means3D", "colors", "opacity", "scales", "rotations", "scale_modifier", "cov3D_precomp", "viewmatrix",
                   "projmatrix", "tan_fovx", "tan_fovy", "image_height", "image_width", "sh", "degree", "campos", "prefiltered", "debug"]
    bwd = list(inspect.signature(dgr._C.rasterize_gaussians_backward).parameters)
    assert bwd == ["background", "means3D", "radii", "colors", "scales", "rotations", "scale_modifier", "cov3D_precomp", "viewmatrix",
                   "projmatrix", "tan_fovx", "tan_fovy", "dL_dout_color", "sh", "degree", "campos", "geomBuffer", "R", "binningBuffer",
                   "imageBuffer", "debug"]
    assert list(inspect.signature(dgr._C.mark_visible).parameters) == ["means3D", "viewmatrix", "projmatrix"]


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The header is plain C: compile a probe with gcc that prints sizeof / offsetof of the two input structs and compare with
    the ctypes mirrors in manus_b200/_lib.py (field names, order, offsets, total size)."""
    import shutil
    import subprocess

    from manus_b200 import _lib

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    pairs = (("mb_raster_inputs", _lib.RasterInputs), ("mb_pose_inputs", _lib.PoseInputs), ("mb_view_inputs", _lib.ViewInputs))
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "manus_b200.h"', "int main(void) {"]
    for cname, cls in pairs:
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for field, _ in cls._fields_:
            lines.append(f'  printf("{cname} {field} %zu\\n", offsetof({cname}, {field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    want = []
    for cname, cls in pairs:
        want.append(f"{cname} size {ctypes.sizeof(cls)}")
        want += [f"{cname} {field} {getattr(cls, field).offset}" for field, _ in cls._fields_]
    assert [g for g in got if g] == want
    # and the header declares no field the mirrors lack
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "manus_b200.h")).read(), flags=re.S)
    for cname, cls in pairs:
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), text, flags=re.S).group(1)
        names = re.findall(r"(\w+)\s*(?=[;,])", body)
        assert sorted(set(names)) == sorted(f for f, _ in cls._fields_), (cname, names)


def test_fused_render_refuses_cpu_tensors(built_lib):
    """There is no CPU path: the fused pose + rasterizer node raises on CPU tensors instead of computing anything."""
    import types

    import torch

    from manus_b200 import _lib
    from manus_b200.render import render_fused

    cam = types.SimpleNamespace(camera_center=torch.zeros(3), height=16, width=16, fovx=0.5, fovy=0.5,
                                world_view_transform=torch.eye(4), full_proj_transform=torch.eye(4))
    params = [torch.zeros(4, 3, requires_grad=True), torch.zeros(4, 3), torch.zeros(4, 4), torch.zeros(4, 1), torch.zeros(4, 1, 3),
              torch.zeros(4, 15, 3)]
    for fused in (True, False):
        with pytest.raises(_lib.ManusB200Error, match="no CPU path"):
            render_fused(params, None, None, cam, torch.ones(3), fuse_backward=fused)
