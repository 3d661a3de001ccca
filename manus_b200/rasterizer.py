"""Drop-in host side of the tile rasterizer: ``GaussianRasterizationSettings`` / ``GaussianRasterizer``.

Same names, argument meaning, return values and error behaviour as the ``diff_gaussian_rasterization`` package MANUS
imports (/root/reference/src/utils/gaussian_utils.py:18-21) and calls (:378-416): 12-field NamedTuple settings,
``GaussianRasterizer(raster_settings)(means3D, means2D, opacities, shs, colors_precomp, scales, rotations,
cov3D_precomp) -> (color[3,H,W], radii[N] int32)``, gradients for (means3D, means2D, sh, colors_precomp, opacities,
scales, rotations, cov3Ds_precomp) with ``means2D`` receiving dL/d(NDC-scaled screen xy) as [N,3] (z = 0) -- the
side channel ``add_densification_stats`` reads (src/models/gaussian.py:335-338).  SURVEY.md section 8b.

All arithmetic happens in the CUDA library behind the C ABI (include/manus_b200.h); torch is used for memory and
the current stream only.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, NamedTuple, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import ptr


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _dense(t: Optional[torch.Tensor], shape=None) -> Optional[torch.Tensor]:
    """fp32, contiguous, on the GPU; None / empty -> None.  Callers hand over slices and [1,4,4] / [1,3] camera tensors."""
    if t is None or t.numel() == 0:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    return t if shape is None else t.reshape(shape)


class CapacityPlan:
    """How the instance capacity of a frame is chosen, and the state that choice needs.  One plan per DEVICE by default
    (``plan_for``); a renderer may hold its own (``SceneRenderer(plan=...)``) so that two scenes on one GPU do not share
    high-water marks.
    'exact': one host read of num_rendered per forward (what upstream does).
    'reserve': no host synchronisation -- capacity is the high-water mark of earlier frames of the same (N, H, W) times a
    margin; an overflow is recorded in the frame's device counters and raised at the next host read (``check_overflow``)."""

    def __init__(self, mode: str = "exact", margin: float = 1.3):
        self.high_water: Dict[tuple, int] = {}
        self.last_state = None      # RasterState of the most recent forward sized by this plan (for check_overflow)
        self.set_mode(mode, margin)

    def set_mode(self, mode: str, margin: float = 1.3) -> None:
        assert mode in ("exact", "reserve")
        self.mode, self.margin = mode, margin
        self.high_water.clear()

    def reserve(self, num_points: int, height: int, width: int, instances: int) -> None:
        key = (int(num_points), int(height), int(width))
        self.high_water[key] = max(self.high_water.get(key, 0), int(instances))


_plans: Dict[int, CapacityPlan] = {}
_plan_defaults = ["exact", 1.3]


def _device_index(device) -> int:
    if device is None:
        return torch.cuda.current_device()
    if isinstance(device, int):
        return device
    device = torch.device(device)
    return torch.cuda.current_device() if device.index is None else device.index


def plan_for(device=None) -> CapacityPlan:
    """The default plan of a device (index, torch.device or None = the current device)."""
    idx = _device_index(device)
    if idx not in _plans:
        _plans[idx] = CapacityPlan(*_plan_defaults)
    return _plans[idx]


def set_capacity_mode(mode: str, margin: float = 1.3, device=None) -> None:
    """device=None: every device's default plan (and the default of plans created later); else that device's plan only."""
    assert mode in ("exact", "reserve")
    if device is None:
        _plan_defaults[:] = [mode, margin]
        for pl in _plans.values():
            pl.set_mode(mode, margin)
    else:
        plan_for(device).set_mode(mode, margin)


def reserve_capacity(device_index: int, num_points: int, height: int, width: int, instances: int) -> None:
    """Reserve-mode sizing: frames of this (device, N, H, W) get room for ``instances * margin`` instances."""
    plan_for(device_index).reserve(num_points, height, width, instances)


def check_overflow(st: "Optional[RasterState]" = None, device=None) -> int:
    """Reserve mode reads nothing back per frame; call this when convenient (it synchronises the stream): returns
    num_rendered of the given forward / the most recent forward of the device's default plan, raises if that frame needed
    more instances than were reserved."""
    st = st if st is not None else plan_for(device).last_state
    if st is None:
        return 0
    st.num_rendered = -1
    st.host_count = None
    return st.resolve()


class RasterState:
    """Opaque state kept from forward to backward (upstream: geomBuffer / binningBuffer / imgBuffer / num_rendered)."""
    __slots__ = ("inputs", "keep", "geom", "binning", "image", "capacity", "num_rendered", "radii", "host_count", "event", "key", "plan")

    def resolve(self) -> int:
        """num_rendered of this frame (waits for the count copy if it is still in flight); raises on overflow."""
        if self.num_rendered < 0:
            if self.host_count is None:      # reserve mode: read the device counters (synchronises the stream)
                self.num_rendered = raster_query(self)[0]
            else:
                self.event.synchronize()
                self.num_rendered = int(self.host_count.item())
            self.plan.high_water[self.key] = max(self.plan.high_water.get(self.key, 0), self.num_rendered)
            if self.num_rendered > self.capacity:
                raise _lib.ManusB200Error(
                    f"instance capacity overflow: frame needed {self.num_rendered} instances, {self.capacity} were reserved; "
                    "the image of this frame is incomplete. Use set_capacity_mode('exact') or a larger margin.")
        return self.num_rendered


def _make_inputs(settings: GaussianRasterizationSettings, means3D, opacities, colors_precomp, shs, cov3D_precomp, scales,
                 rotations, keep: list, num_points: Optional[int] = None, dev=None) -> _lib.RasterInputs:
    dev = means3D.device if dev is None else dev
    bg = _dense(settings.bg.to(dev), (3,))
    view = _dense(settings.viewmatrix.to(dev), (16,))
    proj = _dense(settings.projmatrix.to(dev), (16,))
    cam = _dense(settings.campos.to(dev), (3,))
    keep += [bg, view, proj, cam]
    ri = _lib.RasterInputs()
    ri.num_points = means3D.shape[0] if num_points is None else int(num_points)
    ri.image_width, ri.image_height = int(settings.image_width), int(settings.image_height)
    ri.sh_degree = int(settings.sh_degree)
    ri.sh_coeffs = 0 if shs is None else int(shs.shape[1])
    ri.prefiltered, ri.debug = int(bool(settings.prefiltered)), int(bool(settings.debug))
    tx, ty = settings.tanfovx, settings.tanfovy
    if torch.is_tensor(tx) and tx.is_cuda:
        # camera intrinsics stay on the device (replayable frames: CUDA graphs, no host read of the tensor)
        tx, ty = tx.reshape(-1)[:1], ty.reshape(-1)[:1]
        if not (tx.dtype == ty.dtype == torch.float32 and ty.data_ptr() == tx.data_ptr() + 4):
            tx = torch.cat([tx.float(), ty.float().to(dev)])     # (tanfovx, tanfovy) adjacent in memory
        keep.append(tx)
        ri.tanfovx = ri.tanfovy = 0.0
        ri.tanfov_dev = ptr(tx)
    else:
        ri.tanfovx, ri.tanfovy = float(tx), float(ty)
        ri.tanfov_dev = None
    ri.scale_modifier = float(settings.scale_modifier)
    ri.background, ri.viewmatrix, ri.projmatrix, ri.campos = ptr(bg), ptr(view), ptr(proj), ptr(cam)
    ri.means3D, ri.opacities = ptr(means3D), ptr(opacities)
    ri.colors_precomp, ri.shs = ptr(colors_precomp), ptr(shs)
    ri.cov3D_precomp, ri.scales, ri.rotations = ptr(cov3D_precomp), ptr(scales), ptr(rotations)
    return ri


def rasterize_forward(settings: GaussianRasterizationSettings, means3D, opacities, colors_precomp=None, shs=None,
                      cov3D_precomp=None, scales=None, rotations=None, capacity: Optional[int] = None, exact: bool = False,
                      pose_inputs=None, plan: Optional[CapacityPlan] = None):
    """-> (color[3,H,W], radii[N], RasterState).  Inputs must already be dense fp32 CUDA tensors (or None).
    exact=True: size the instance buffers from this frame's own num_rendered (one host read) whatever the capacity mode.
    pose_inputs: a filled ``_lib.PoseInputs`` -- the per-Gaussian stage then is mb_pose_project_forward (LBS + covariance +
    SH->RGB + projection in ONE kernel) instead of mb_raster_forward_geom on precomputed arrays; means3D / opacities /
    colors_precomp / cov3D_precomp are then OUTPUT buffers of that kernel, or all None (the posed arrays never touch HBM;
    only the fused backward, mb_pose_backward_from_raster, can follow).
    plan: the CapacityPlan that sizes the instance buffers (default: the device's, ``plan_for``)."""
    L = _lib.lib()
    if pose_inputs is None and not means3D.is_cuda:
        raise _lib.ManusB200Error("manus_b200 rasterizer needs CUDA tensors (there is no CPU path)")
    if pose_inputs is not None:
        N = int(pose_inputs.num_points)
        dev = settings.viewmatrix.device
    else:
        N, dev = means3D.shape[0], means3D.device
    H, W = int(settings.image_height), int(settings.image_width)
    st = RasterState()
    st.keep = [means3D, opacities, colors_precomp, shs, cov3D_precomp, scales, rotations]
    st.inputs = _make_inputs(settings, means3D, opacities, colors_precomp, shs, cov3D_precomp, scales, rotations, st.keep, N, dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        st.geom = torch.empty(L.mb_raster_geom_bytes(N), dtype=torch.uint8, device=dev)
        st.image = torch.empty(L.mb_raster_image_bytes(W, H), dtype=torch.uint8, device=dev)
        st.radii = torch.empty(N, dtype=torch.int32, device=dev)
        color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        st.plan = plan if plan is not None else plan_for(dev)
        st.key = (N, H, W)
        reserve = capacity is None and not exact and st.plan.mode == "reserve" and st.key in st.plan.high_water
        # reserve mode: nothing is read back per frame (the frame can be captured in a CUDA graph); an overflow is
        # recorded in the device counters and raised by check_overflow() / raster_query()
        st.host_count = None if reserve else torch.zeros(1, dtype=torch.int64).pin_memory()
        count_ptr = None if reserve else st.host_count.data_ptr()
        if pose_inputs is not None:
            _lib.check(L.mb_pose_project_forward(C.byref(pose_inputs), C.byref(st.inputs), ptr(st.geom), st.geom.numel(), ptr(st.radii),
                                                 count_ptr, ptr(means3D), ptr(cov3D_precomp), ptr(colors_precomp), ptr(opacities), stream),
                       "mb_pose_project_forward")
        else:
            _lib.check(L.mb_raster_forward_geom(C.byref(st.inputs), ptr(st.geom), st.geom.numel(), ptr(st.radii), count_ptr, stream),
                       "mb_raster_forward_geom")
        st.num_rendered = -1
        st.event = None
        if reserve:
            capacity = int(st.plan.high_water[st.key] * st.plan.margin) + 1024      # no host synchronisation
        else:
            st.event = torch.cuda.Event()
            st.event.record(torch.cuda.current_stream(dev))
            if capacity is None:
                st.capacity = 1 << 62
                capacity = st.resolve()                                          # one 8-byte host read, like upstream
        st.capacity = int(capacity)
        st.binning = torch.empty(L.mb_raster_binning_bytes(st.capacity, W, H), dtype=torch.uint8, device=dev)
        _lib.check(L.mb_raster_forward_render(C.byref(st.inputs), ptr(st.geom), ptr(st.binning), st.binning.numel(), st.capacity,
                                              ptr(st.image), st.image.numel(), ptr(color), stream), "mb_raster_forward_render")
    st.plan.last_state = st
    return color, st.radii, st


def raster_query(st: RasterState):
    """(num_rendered, num_visible, overflow) of a finished forward; synchronises the current stream."""
    L = _lib.lib()
    nr, nv, ov = C.c_int64(0), C.c_int64(0), C.c_int32(0)
    dev = st.geom.device
    with torch.cuda.device(dev):
        _lib.check(L.mb_raster_query(ptr(st.geom), C.byref(nr), C.byref(nv), C.byref(ov),
                                     torch.cuda.current_stream(dev).cuda_stream), "mb_raster_query")
    return nr.value, nv.value, ov.value


def rasterize_backward(st: RasterState, grad_color: torch.Tensor):
    """grad_color: [3,H,W] with ANY strides (the permuted HWC view the MANUS losses produce is read in place)."""
    L = _lib.lib()
    ri = st.inputs
    N, M = ri.num_points, ri.sh_coeffs
    dev = st.geom.device
    if grad_color.dtype != torch.float32:
        grad_color = grad_color.float()
    sc, sy, sx = grad_color.stride()
    new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    g_means2D, g_colors, g_opacity, g_means3D, g_cov3D = new(N, 3), new(N, 3), new(N, 1), new(N, 3), new(N, 6)
    g_sh = new(N, M, 3) if ri.shs else None
    g_scales = new(N, 3) if not ri.cov3D_precomp else None
    g_rot = new(N, 4) if not ri.cov3D_precomp else None
    with torch.cuda.device(dev):
        scratch = torch.empty(L.mb_raster_backward_scratch_bytes(N), dtype=torch.uint8, device=dev)
        _lib.check(L.mb_raster_backward(C.byref(ri), ptr(st.radii), ptr(st.geom), ptr(st.binning), st.capacity, ptr(st.image),
                                        ptr(grad_color), sc, sy, sx, ptr(scratch), scratch.numel(), ptr(g_means2D),
                                        ptr(g_colors), ptr(g_opacity), ptr(g_means3D), ptr(g_cov3D), ptr(g_sh), ptr(g_scales),
                                        ptr(g_rot), torch.cuda.current_stream(dev).cuda_stream), "mb_raster_backward")
    return g_means2D, g_colors, g_opacity, g_means3D, g_cov3D, g_sh, g_scales, g_rot


def rasterize_backward_blend(st: RasterState, grad_color: torch.Tensor) -> torch.Tensor:
    """The tile half of the backward only (mb_raster_backward_blend): returns the accumulator rows ([N,12] fp32, as bytes)
    that ``manus_b200.pose.pose_backward_from_raster`` consumes.  grad_color: [3,H,W] with any strides."""
    L = _lib.lib()
    ri = st.inputs
    dev = st.geom.device
    if grad_color.dtype != torch.float32:
        grad_color = grad_color.float()
    sc, sy, sx = grad_color.stride()
    with torch.cuda.device(dev):
        scratch = torch.empty(L.mb_raster_backward_scratch_bytes(ri.num_points), dtype=torch.uint8, device=dev)
        _lib.check(L.mb_raster_backward_blend(C.byref(ri), ptr(st.radii), ptr(st.geom), ptr(st.binning), st.capacity, ptr(st.image),
                                              ptr(grad_color), sc, sy, sx, ptr(scratch), scratch.numel(),
                                              torch.cuda.current_stream(dev).cuda_stream), "mb_raster_backward_blend")
    return scratch


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
        d = lambda t: _dense(t)
        m3, op = d(means3D), d(opacities)
        if m3 is None:   # N == 0: upstream returns early with the background image
            H, W = int(raster_settings.image_height), int(raster_settings.image_width)
            dev = means3D.device
            color = raster_settings.bg.to(dev).float().reshape(3, 1, 1).expand(3, H, W).contiguous()
            ctx.state = None
            return color, torch.zeros(0, dtype=torch.int32, device=dev)
        if m3.dim() != 2 or m3.shape[1] != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        color, radii, st = rasterize_forward(raster_settings, m3, op.reshape(-1), d(colors_precomp), d(sh), d(cov3Ds_precomp),
                                             d(scales), d(rotations))
        ctx.state = st
        ctx.shapes = (means3D.shape, means2D.shape, opacities.shape)
        ctx.mark_non_differentiable(radii)
        ctx.set_materialize_grads(False)      # no zero tensor for the (integer) radii output on every backward
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii):
        st = ctx.state
        if st is None or grad_out_color is None:
            return (None,) * 9
        if st.host_count is not None:
            st.resolve()
        g_means2D, g_colors, g_opacity, g_means3D, g_cov3D, g_sh, g_scales, g_rot = rasterize_backward(st, grad_out_color)
        m3_shape, m2_shape, op_shape = ctx.shapes
        ctx.state = None
        return (g_means3D.reshape(m3_shape), g_means2D.reshape(m2_shape), g_sh, g_colors if not st.inputs.shs else None,
                g_opacity.reshape(op_shape), g_scales, g_rot, g_cov3D if st.inputs.cov3D_precomp else None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                     raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        L = _lib.lib()
        with torch.no_grad():
            p = _dense(positions)
            n = 0 if p is None else p.shape[0]
            out = torch.zeros(n, dtype=torch.uint8, device=positions.device)
            if n:
                rs = self.raster_settings
                view, proj = _dense(rs.viewmatrix.to(p.device), (16,)), _dense(rs.projmatrix.to(p.device), (16,))
                with torch.cuda.device(p.device):
                    _lib.check(L.mb_mark_visible(ptr(p), n, ptr(view), ptr(proj), ptr(out),
                                                 torch.cuda.current_stream(p.device).cuda_stream), "mb_mark_visible")
            return out.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        empty = torch.Tensor([])
        return rasterize_gaussians(means3D, means2D, empty if shs is None else shs,
                                   empty if colors_precomp is None else colors_precomp, opacities,
                                   empty if scales is None else scales, empty if rotations is None else rotations,
                                   empty if cov3D_precomp is None else cov3D_precomp, self.raster_settings)


# ---- the function-level surface of upstream's pybind11 module ``diff_gaussian_rasterization._C`` ----

def _state_from_buffers(settings, means3D, colors, sh, cov3D_precomp, scales, rotations, radii, geom, binning, image, num_rendered):
    d = _dense
    st = RasterState()
    m3 = d(means3D)
    st.keep = [m3, d(colors), d(sh), d(cov3D_precomp), d(scales), d(rotations)]
    st.inputs = _make_inputs(settings, m3, None, st.keep[1], st.keep[2], st.keep[3], st.keep[4], st.keep[5], st.keep)
    st.geom, st.binning, st.image, st.radii = geom, binning, image, radii
    st.capacity = st.num_rendered = int(num_rendered)
    st.host_count, st.event, st.key = None, None, None
    return st


def _c_settings(bg, scale_modifier, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, degree, campos, debug):
    return GaussianRasterizationSettings(int(image_height), int(image_width), float(tan_fovx), float(tan_fovy), bg, float(scale_modifier),
                                         viewmatrix, projmatrix, int(degree), campos, False, bool(debug))


def c_rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                          projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos, prefiltered, debug):
    """``_C.rasterize_gaussians`` with upstream's positional signature and return value
    (num_rendered, out_color[3,H,W], radii[N] int32, geomBuffer u8, binningBuffer u8, imgBuffer u8); absent tensors are passed
    as empty tensors like upstream's Python wrapper does.  The three byte buffers are this library's opaque state."""
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    settings = _c_settings(background, scale_modifier, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, degree,
                           campos, debug)
    d = _dense
    if means3D.shape[0] == 0:
        dev = means3D.device
        u8 = torch.empty(0, dtype=torch.uint8, device=dev)
        color = background.to(dev).float().reshape(3, 1, 1).expand(3, int(image_height), int(image_width)).contiguous()
        return 0, color, torch.zeros(0, dtype=torch.int32, device=dev), u8, u8.clone(), u8.clone()
    color, radii, st = rasterize_forward(settings, d(means3D), d(opacity).reshape(-1), d(colors), d(sh), d(cov3D_precomp), d(scales),
                                         d(rotations), exact=True)      # capacity == num_rendered: the backward re-derives the layout from R
    return st.resolve(), color, radii, st.geom, st.binning, st.image


def c_rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                                   projmatrix, tan_fovx, tan_fovy, dL_dout_color, sh, degree, campos, geomBuffer, R, binningBuffer,
                                   imageBuffer, debug):
    """``_C.rasterize_gaussians_backward`` with upstream's positional signature and its 8-tuple
    (dL_dmeans2D[N,3], dL_dcolors[N,3], dL_dopacity[N,1], dL_dmeans3D[N,3], dL_dcov3D[N,6], dL_dsh[N,M,3], dL_dscales[N,3],
    dL_drotations[N,4]); gradients of absent inputs come back as zeros of upstream's shapes."""
    N, dev = means3D.shape[0], means3D.device
    H, W = int(dL_dout_color.shape[-2]), int(dL_dout_color.shape[-1])
    M = 0 if sh is None or sh.numel() == 0 else int(sh.shape[1])
    z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
    if N == 0 or R == 0 and geomBuffer.numel() == 0:
        return z(N, 3), z(N, 3), z(N, 1), z(N, 3), z(N, 6), z(N, M, 3), z(N, 3), z(N, 4)
    settings = _c_settings(background, scale_modifier, viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, degree, campos, debug)
    st = _state_from_buffers(settings, means3D, colors, sh, cov3D_precomp, scales, rotations, radii, geomBuffer, binningBuffer,
                             imageBuffer, R)
    g2d, gcol, gop, g3d, gcov, gsh, gsc, grot = rasterize_backward(st, dL_dout_color)
    return (g2d, gcol, gop.reshape(N, 1), g3d, gcov, gsh if gsh is not None else z(N, M, 3), gsc if gsc is not None else z(N, 3),
            grot if grot is not None else z(N, 4))


def c_mark_visible(means3D, viewmatrix, projmatrix):
    """``_C.mark_visible(means3D, viewmatrix, projmatrix) -> bool[N]``."""
    rs = GaussianRasterizationSettings(0, 0, 1.0, 1.0, None, 1.0, viewmatrix, projmatrix, 0, None, False, False)
    return GaussianRasterizer(rs).markVisible(means3D)


def debug_views(st: RasterState):
    """Typed views into the opaque buffers of a finished forward (offsets from mb_raster_state_layout).
    For tests and diagnostics only: {'final_T' [H,W], 'n_contrib' [H,W], 'ranges' [tiles,2], 'tile_maxlast' [tiles],
    'point_list' [num_rendered] (gaussian id per sorted instance), 'records' [N,12]}."""
    L = _lib.lib()
    H, W, N = st.inputs.image_height, st.inputs.image_width, st.inputs.num_points
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    px = H * W
    off = (C.c_int64 * 8)()
    _lib.check(L.mb_raster_state_layout(N, st.capacity, W, H, off, 8), "mb_raster_state_layout")
    img = st.image
    final_T = img[off[0]: off[0] + px * 4].view(torch.float32).reshape(H, W)
    n_contrib = img[off[1]: off[1] + px * 4].view(torch.int32).reshape(H, W)
    ranges = img[off[2]: off[2] + tiles * 8].view(torch.int32).reshape(tiles, 2)
    maxlast = img[off[3]: off[3] + tiles * 4].view(torch.int32)
    n = st.resolve()
    point_list = st.binning[off[4]: off[4] + min(n, st.capacity) * 4].view(torch.int32)
    records = st.geom[off[6]: off[6] + N * 48].view(torch.float32).reshape(N, 12)
    return dict(final_T=final_T, n_contrib=n_contrib, ranges=ranges, tile_maxlast=maxlast, point_list=point_list, records=records)
