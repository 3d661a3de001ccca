"""``render_gaussians``: the caller of the rasterizer boundary, same signature and return dict as
/root/reference/src/utils/gaussian_utils.py:349-428, plus the fused entry ``render_fused`` that starts from the raw
GaussianModel parameters (one pose kernel + the rasterizer; nothing else runs per frame).
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from ._lib import ptr
from .pose import _f32c, _inputs, pose_gaussians
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, rasterize_backward_blend, rasterize_forward


def _scalar(v) -> float:
    return float(v.reshape(-1)[0]) if torch.is_tensor(v) else float(v)


def _settings(camera, bg_color, sh_degree, device) -> GaussianRasterizationSettings:
    tan_dev = getattr(camera, "tanfov_dev", None)     # optional CUDA tensor (tanfovx, tanfovy): intrinsics stay on the device
    return GaussianRasterizationSettings(
        image_height=int(_scalar(camera.height)), image_width=int(_scalar(camera.width)),
        tanfovx=tan_dev[0:1] if tan_dev is not None else math.tan(_scalar(camera.fovx) * 0.5),
        tanfovy=tan_dev[1:2] if tan_dev is not None else math.tan(_scalar(camera.fovy) * 0.5),
        bg=bg_color, scale_modifier=1,
        viewmatrix=torch.as_tensor(camera.world_view_transform).to(device),
        projmatrix=torch.as_tensor(camera.full_proj_transform).to(device),
        sh_degree=sh_degree, campos=torch.as_tensor(camera.camera_center).to(device), prefiltered=False, debug=False)


_SH_C0 = 0.28209479177387814
_SH_C1 = 0.4886025119029199
_SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435)


def eval_sh(deg: int, sh: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """src/utils/sh_utils.py:57-120 for degrees 0..3: sh [N,3,K], unit dirs [N,3] -> [N,3] (plain torch, any device)."""
    result = _SH_C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        result = result - _SH_C1 * y * sh[..., 1] + _SH_C1 * z * sh[..., 2] - _SH_C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            result = (result + _SH_C2[0] * xy * sh[..., 4] + _SH_C2[1] * yz * sh[..., 5] + _SH_C2[2] * (2.0 * zz - xx - yy) * sh[..., 6]
                      + _SH_C2[3] * xz * sh[..., 7] + _SH_C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                result = (result + _SH_C3[0] * y * (3 * xx - yy) * sh[..., 9] + _SH_C3[1] * xy * z * sh[..., 10]
                          + _SH_C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + _SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
                          + _SH_C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + _SH_C3[5] * z * (xx - yy) * sh[..., 14]
                          + _SH_C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return result


class _ShColors(torch.autograd.Function):
    """mb_sh_colors_forward / backward: one thread per Gaussian, closed-form 4x4 inverse (csrc/shcolor.cu)."""

    @staticmethod
    def forward(ctx, means, features, tf, campos, sh_degree):
        L = _lib.lib()
        m, f = _f32c(means), _f32c(features)
        t = None if tf is None else _f32c(tf).reshape(-1, 4, 4)
        c = _f32c(campos).reshape(-1)[:3].contiguous()
        N, K = m.shape[0], f.shape[1]
        colors = torch.empty((N, 3), dtype=torch.float32, device=m.device)
        with torch.cuda.device(m.device):
            _lib.check(L.mb_sh_colors_forward(ptr(m), ptr(f), ptr(t), ptr(c), N, int(sh_degree), K, ptr(colors),
                                              torch.cuda.current_stream(m.device).cuda_stream), "mb_sh_colors_forward")
        ctx.save_for_backward(m, f, c, *(() if t is None else (t,)))
        ctx.sh_degree, ctx.has_tf = int(sh_degree), t is not None
        ctx.shapes = (means.shape, features.shape, None if tf is None else tf.shape)
        return colors

    @staticmethod
    def backward(ctx, g_colors):
        L = _lib.lib()
        saved = ctx.saved_tensors
        m, f, c = saved[:3]
        t = saved[3] if ctx.has_tf else None
        N, K = m.shape[0], f.shape[1]
        g = _f32c(g_colors)
        g_m, g_f = torch.empty_like(m), torch.empty_like(f)
        g_t = torch.empty_like(t) if (t is not None and ctx.needs_input_grad[2]) else None
        with torch.cuda.device(m.device):
            _lib.check(L.mb_sh_colors_backward(ptr(m), ptr(f), ptr(t), ptr(c), N, ctx.sh_degree, K, ptr(g), ptr(g_m), ptr(g_f), ptr(g_t),
                                               torch.cuda.current_stream(m.device).cuda_stream), "mb_sh_colors_backward")
        sm, sf, st = ctx.shapes
        return g_m.reshape(sm), g_f.reshape(sf), None if g_t is None else g_t.reshape(st), None, None


def calculate_colors_from_sh(posed_means, cano_features, cano_means, camera, sh_degree, tf):
    """gaussian_utils.py:431-449 for callers that hold the materialised per-Gaussian ``tf`` [N,4,4] (unchanged MANUS): the view
    direction is taken in canonical space through inv(tf).  One kernel forward, one backward (gradients to the features, to the
    means and to tf); the reference's torch.linalg.inv on N 4x4 matrices becomes a closed-form inverse per thread.  The fused
    ``pose_gaussians`` kernel does the same arithmetic without materialising tf and remains the fast path."""
    if not posed_means.is_cuda:
        raise _lib.ManusB200Error("manus_b200.render.calculate_colors_from_sh needs CUDA tensors (there is no CPU path)")
    cc = torch.as_tensor(camera.camera_center).to(posed_means.device)
    means = cano_means if tf is not None else posed_means
    return _ShColors.apply(means, cano_features, tf, cc, int(sh_degree))


_ZEROS = {}


def _screenspace_leaf(posed_means):
    """A [N,3] zero tensor that receives the screen-space gradient, without a fill kernel per frame: the rasterizer never
    reads means2D, so a cached zero buffer is re-used and only a fresh autograd leaf is made on top of it."""
    key = (posed_means.device, tuple(posed_means.shape), posed_means.dtype)
    z = _ZEROS.get(key)
    if z is None:
        _ZEROS.clear()
        z = _ZEROS[key] = torch.zeros_like(posed_means, requires_grad=False)
    return z.detach().requires_grad_(True)


def render_gaussians(posed_means, posed_cov, cano_means, cano_features, cano_opacity, camera, bg_color, colors_precomp=None,
                     sh_degree=3, tf=None, device=torch.device("cuda"), _cached_screenspace=False):
    """gaussian_utils.py:349-428.  With ``colors_precomp=None`` the colours come from the SH features like the reference
    (:401-404): for ``tf=None`` (static object: world-space view direction) the rasterizer kernels evaluate the SH themselves
    (forward and backward), otherwise ``calculate_colors_from_sh`` runs first."""
    if _cached_screenspace:
        screenspace_points = _screenspace_leaf(posed_means)          # a leaf: .grad is populated without retain_grad()
    else:
        screenspace_points = torch.zeros_like(posed_means, dtype=posed_means.dtype, requires_grad=True, device=device) + 0
        try:
            screenspace_points.retain_grad()
        except Exception:
            pass
    shs = None
    if colors_precomp is None:
        if cano_features is None:
            raise ValueError("render_gaussians: pass cano_features (SH) or colors_precomp")
        if tf is None:
            shs = cano_features[:, : (sh_degree + 1) ** 2]         # in-kernel SH -> RGB along (mean - camera centre)
        else:
            colors_precomp = calculate_colors_from_sh(posed_means, cano_features, cano_means, camera, sh_degree, tf)
    rasterizer = GaussianRasterizer(raster_settings=_settings(camera, bg_color, sh_degree, device))
    rendered_image, radii = rasterizer(means3D=posed_means, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp,
                                       opacities=cano_opacity, scales=None, rotations=None, cov3D_precomp=posed_cov)
    rendered_image = torch.permute(rendered_image, (1, 2, 0))
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii}


class DeferredPoseBackward:
    """The second half of the fused backward of one view -- mb_pose_backward_from_raster on the blend backward's accumulator
    rows -- kept for the caller to launch on ranges of Gaussians (``run(lo, hi, accumulate)``; lo must be a multiple of 128,
    the pose kernels' tile, so that every sub-array stays 16-byte aligned for the bulk copies)."""

    def __init__(self, t, cam, sh_degree, isotropic, num_skinned, st, scratch, g_means2D, grads, stats=None):
        self.t, self.cam, self.sh_degree, self.isotropic, self.num_skinned = t, cam, sh_degree, isotropic, num_skinned
        self.st, self.scratch, self.g_means2D, self.grads, self.stats = st, scratch, g_means2D, grads, stats

    def run(self, lo: int, hi: int, accumulate: bool) -> None:
        L = _lib.lib()
        t, N = self.t, self.t[0].shape[0]
        if lo % 128 or not (0 <= lo < hi <= N):
            raise ValueError(f"bad Gaussian range [{lo}, {hi}) of {N} (lo must be a multiple of 128)")
        n, ns = hi - lo, max(0, min(self.num_skinned, hi) - lo)
        row = lambda v: None if v is None else v[lo:hi]
        skin = None if (t[6] is None or ns == 0) else t[6][lo:lo + ns]
        tt = [row(v) for v in t[:6]] + [skin, t[7]]
        pi = _inputs(*tt, self.cam, self.sh_degree, self.isotropic, ns)
        if skin is None and t[6] is not None:
            pi.num_bones = t[6].shape[1]
        ri = type(self.st.inputs)()
        C.memmove(C.byref(ri), C.byref(self.st.inputs), C.sizeof(ri))
        ri.num_points = n
        g = [row(v) for v in self.grads]
        dev = t[0].device
        with torch.cuda.device(dev):
            _lib.check(L.mb_pose_backward_from_raster(C.byref(pi), C.byref(ri), ptr(self.st.radii[lo:hi]), self.scratch.data_ptr() + lo * 48,
                                                      ptr(self.g_means2D[lo:hi]), ptr(g[0]), ptr(g[1]), ptr(g[2]), ptr(g[3]), ptr(g[4]),
                                                      ptr(g[5]) if g[5] is not None and g[5].numel() else None, None,
                                                      int(accumulate), *_stat_ptrs(self.stats, N, lo),
                                                      torch.cuda.current_stream(dev).cuda_stream),
                       "mb_pose_backward_from_raster")


def pose_backward_views(defs, lo: int, hi: int, accumulate: bool) -> None:
    """The deferred pose backwards of ALL the views of a step over Gaussians [lo, hi) in ONE kernel pass
    (mb_pose_backward_from_raster_views): the parameters are staged once per tile, every thread walks the views of its Gaussian.
    ``defs``: the DeferredPoseBackward objects of the step's views (same parameters, same gradient sink)."""
    L = _lib.lib()
    d0 = defs[0]
    t, N = d0.t, d0.t[0].shape[0]
    if lo % 128 or not (0 <= lo < hi <= N):
        raise ValueError(f"bad Gaussian range [{lo}, {hi}) of {N} (lo must be a multiple of 128)")
    n, ns = hi - lo, max(0, min(d0.num_skinned, hi) - lo)
    row = lambda v: None if v is None else v[lo:hi]
    skin = None if (t[6] is None or ns == 0) else t[6][lo:lo + ns]
    pi = _inputs(*([row(v) for v in t[:6]] + [skin, t[7]]), d0.cam, d0.sh_degree, d0.isotropic, ns)
    if skin is None and t[6] is not None:
        pi.num_bones = t[6].shape[1]
    views = (_lib.ViewInputs * len(defs))()
    keep = []
    for k, d in enumerate(defs):
        if d.t[0].data_ptr() != t[0].data_ptr() or d.grads[0].data_ptr() != d0.grads[0].data_ptr():
            raise RuntimeError("the views of a step must share their parameters and their gradient sink")
        ri = type(d.st.inputs)()
        C.memmove(C.byref(ri), C.byref(d.st.inputs), C.sizeof(ri))
        ri.num_points = n
        keep.append(ri)
        w = views[k]
        w.raster = C.addressof(ri)
        w.radii = d.st.radii.data_ptr() + 4 * lo
        w.grad_scratch = d.scratch.data_ptr() + 48 * lo
        w.dL_dmeans2D = d.g_means2D.data_ptr() + 12 * lo
        bt = d.t[7]
        if isinstance(bt, (tuple, list)):
            w.bone_tf, w.bones_posed = None, bt[0].data_ptr()
        else:
            w.bone_tf, w.bones_posed = ptr(bt), None
        w.campos = d.cam.data_ptr()
    g = [row(v) for v in d0.grads]
    dev = t[0].device
    with torch.cuda.device(dev):
        _lib.check(L.mb_pose_backward_from_raster_views(C.byref(pi), len(defs), views, ptr(g[0]), ptr(g[1]), ptr(g[2]), ptr(g[3]), ptr(g[4]),
                                                        ptr(g[5]) if g[5] is not None and g[5].numel() else None, int(accumulate),
                                                        *_stat_ptrs(d0.stats, N, lo), torch.cuda.current_stream(dev).cuda_stream),
                   "mb_pose_backward_from_raster_views")


def _stat_ptrs(stats, n: int, lo: int = 0):
    """(xyz_gradient_accum, denom, max_radii2D) -> the three device pointers (offset by ``lo`` Gaussians), or three NULLs."""
    if stats is None:
        return None, None, None
    out = []
    for s in stats:
        if not (s.is_cuda and s.dtype == torch.float32 and s.is_contiguous() and s.numel() == n):
            raise RuntimeError("densification statistics must be dense fp32 CUDA tensors with one element per Gaussian")
        out.append(s.data_ptr() + 4 * lo)
    return tuple(out)


class _RenderFused(torch.autograd.Function):
    """pose forward + rasterizer forward as ONE autograd node, so that the backward can keep the rasterizer's per-Gaussian
    gradients out of HBM: tile backward (accumulator rows) -> pose backward with the projection backward inside
    (mb_pose_backward_from_raster).  Same results as pose_gaussians + GaussianRasterizer up to fp32 contraction order."""

    @staticmethod
    def forward(ctx, xyz, log_scale, quat, opacity_logit, f_dc, f_rest, skin_wts, screenspace, bone_tf, campos, settings, sh_degree,
                isotropic, num_skinned, grad_sink, accumulate, want_posed, plan=None):
        L = _lib.lib()
        if not xyz.is_cuda:
            raise _lib.ManusB200Error("manus_b200.render_fused needs CUDA tensors (there is no CPU path)")
        dev = xyz.device
        t = [_f32c(v) for v in (xyz, log_scale, quat, opacity_logit, f_dc, f_rest, skin_wts)]
        if isinstance(bone_tf, (tuple, list)):      # (bones_posed, bones_rest_inv, B): transforms built inside the kernels
            posed_b, rest_inv = _f32c(bone_tf[0]).reshape(-1, 4, 4), _f32c(bone_tf[1]).reshape(-1, 4, 4)
            n_rows = int(bone_tf[2]) if len(bone_tf) > 2 else rest_inv.shape[0]
            if posed_b.shape != rest_inv.shape or n_rows < rest_inv.shape[0]:
                raise RuntimeError(f"bones_posed {tuple(posed_b.shape)} / bones_rest_inv {tuple(rest_inv.shape)} / B={n_rows} do not match")
            t.append((posed_b, rest_inv))
        else:
            t.append(_f32c(bone_tf))
            n_rows = None if t[7] is None else t[7].shape[0]
        cam = _f32c(campos).reshape(-1)[:3].contiguous()
        N = t[0].shape[0]
        if t[6] is not None and (t[6].shape[0] != num_skinned or t[7] is None or n_rows != t[6].shape[1]):
            raise RuntimeError(f"skin_wts {tuple(t[6].shape)} does not match num_skinned={num_skinned} / {n_rows} bone transforms")   # hand_dynamic.py:104
        pi = _inputs(*t, cam, sh_degree, isotropic, num_skinned)
        new = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        if want_posed:
            posed_xyz, cov6, colors, opacity = new(N, 3), new(N, 6), new(N, 3), new(N, 1)
        else:       # the posed arrays stay in registers between the pose step and the projection
            posed_xyz = cov6 = colors = opacity = None
        # ONE kernel for LBS + covariance + SH->RGB + projection (mb_pose_project_forward), then binning + tile kernels
        color, radii, st = rasterize_forward(settings, posed_xyz, None if opacity is None else opacity.reshape(-1),
                                             colors_precomp=colors, cov3D_precomp=cov6, pose_inputs=pi, plan=plan)
        st.keep += [v for v in t[:7]] + list(t[7] if isinstance(t[7], tuple) else [t[7]]) + [cam]
        ctx.saved = (t, cam, sh_degree, isotropic, num_skinned, st)
        ctx.grad_sink, ctx.accumulate = grad_sink, bool(accumulate) and grad_sink is not None
        ctx.need_skin = skin_wts is not None and skin_wts.requires_grad
        ctx.shapes = [None if v is None else v.shape for v in (xyz, log_scale, quat, opacity_logit, f_dc, f_rest, skin_wts, screenspace)]
        ctx.set_materialize_grads(False)
        if not want_posed:
            ctx.mark_non_differentiable(radii)
            return color, radii, None, None, None, None
        ctx.mark_non_differentiable(radii, posed_xyz, cov6, colors, opacity)
        return color, radii, posed_xyz, cov6, colors, opacity

    @staticmethod
    def backward(ctx, g_color, *_unused):
        L = _lib.lib()
        t, cam, sh_degree, isotropic, num_skinned, st = ctx.saved
        if any(g is not None for g in _unused):
            # posed_xyz / posed_cov / colors / cano_opacity are outputs for inspection (contact maps, logging): this node does not
            # propagate gradients that arrive through them -- say so instead of dropping them
            raise RuntimeError("render_fused: a loss term depends on posed_xyz / posed_cov / colors / cano_opacity; use "
                               "fuse_backward=False (pose_gaussians + render_gaussians) to differentiate through them")
        if g_color is None:
            return (None,) * 18
        if st.host_count is not None:
            st.resolve()
        dev = t[0].device
        N = t[0].shape[0]
        scratch = rasterize_backward_blend(st, g_color)
        pi = _inputs(*t, cam, sh_degree, isotropic, num_skinned)
        sink = ctx.grad_sink
        if sink is not None:
            g = [sink.get(k) for k in ("xyz", "log_scale", "quat", "opacity_logit", "f_dc", "f_rest")]
            for k, (gk, ref) in enumerate(zip(g, t[:6])):
                if gk is None and k == 5:
                    continue                     # compact exchange: the f_rest gradient is rebuilt from g_f_dc
                if ref is not None and (gk is None or gk.numel() != ref.numel() or not gk.is_contiguous() or gk.dtype != torch.float32):
                    raise RuntimeError("grad_sink tensors must be dense fp32 with the parameter's size")
        else:
            g = [None if ref is None else torch.empty_like(ref) for ref in t[:6]]
        g_skin = None
        if ctx.need_skin:
            g_skin = torch.zeros_like(t[6]) if ctx.accumulate else torch.empty_like(t[6])
        g_means2D = torch.empty((N, 3), dtype=torch.float32, device=dev)
        if sink is not None and sink.get("_defer") is not None:
            # the caller runs the pose backward itself, range by range (DeferredPoseBackward.run): the data-parallel step
            # all-reduces the gradients of one range of Gaussians while the pose backward of the next range computes
            if ctx.need_skin:
                raise RuntimeError("a deferred pose backward does not produce skin-weight gradients")
            sink["_defer"].append(DeferredPoseBackward(t, cam, sh_degree, isotropic, num_skinned, st, scratch, g_means2D, g, sink.get("_stats")))
        else:
            with torch.cuda.device(dev):
                if sink is not None and sink.get("_wait") is not None:
                    torch.cuda.current_stream(dev).wait_event(sink["_wait"])
                _lib.check(L.mb_pose_backward_from_raster(C.byref(pi), C.byref(st.inputs), ptr(st.radii), ptr(scratch), ptr(g_means2D),
                                                          ptr(g[0]), ptr(g[1]), ptr(g[2]), ptr(g[3]), ptr(g[4]),
                                                          ptr(g[5]) if g[5] is not None and g[5].numel() else None, ptr(g_skin),
                                                          int(ctx.accumulate), *_stat_ptrs(None if sink is None else sink.get("_stats"), N),
                                                          torch.cuda.current_stream(dev).cuda_stream),
                           "mb_pose_backward_from_raster")
                if sink is not None and sink.get("_record") is not None:
                    sink["_record"].record(torch.cuda.current_stream(dev))
        sh = ctx.shapes
        rs = lambda v, s: None if v is None else v.reshape(s)
        ctx.saved = None
        head = (None,) * 6 if sink is not None else tuple(rs(gk, s) for gk, s in zip(g, sh[:6]))
        return head + (rs(g_skin, sh[6]), rs(g_means2D, sh[7])) + (None,) * 10


def render_fused(params, skin_wts, bone_tf, camera, bg_color, sh_degree=3, isotropic=False, num_skinned=None, grad_sink=None,
                 accumulate=False, fuse_backward=False, want_posed=True, plan=None):
    """params: (xyz, log_scale, quat, opacity_logit, f_dc, f_rest) -- the six nn.Parameters of GaussianModel.
    Returns the render_gaussians dict plus the posed quantities (what TrainingModule.forward returns).
    bone_tf: [B,4,4] transforms, or a (bones_posed[nb,4,4], bones_rest_inv[nb,4,4], B) triple -- the kernels then build
    T_b = posed_b rest_b^-1 (+ identity rows up to B) themselves (hand_dynamic.py:93-102).
    want_posed=False (fused backward only): posed_xyz / posed_cov / colors / cano_opacity are not produced (None in the dict);
    they never touch HBM.
    plan: the rasterizer.CapacityPlan that sizes the frame's instance buffers (default: the device's plan).
    grad_sink: {name: dense fp32 tensor} -- the backward writes (accumulate=True: adds) the six parameter gradients of THIS
    node there and returns None for them to autograd.  Only this node's contribution goes to the sink: a gradient that reaches a
    parameter by another route -- skin weights looked up from xyz (skinning_weights_from_voxel_grid, the reference's
    mano_init_voxel mode: its backward adds to xyz.grad through autograd), a regulariser on the leaves -- lands in the leaf's
    .grad as usual and has to be added to the sink by the caller.  With fuse_backward=True the posed outputs are for
    inspection only (a loss term on them raises in the backward)."""
    xyz, log_scale, quat, opacity_logit, f_dc, f_rest = params
    device = xyz.device
    campos = torch.as_tensor(camera.camera_center).to(device)
    if fuse_backward:
        # one autograd node for pose + rasterizer: the backward runs the projection backward inside the pose backward kernel
        if num_skinned is None:
            num_skinned = 0 if skin_wts is None else skin_wts.shape[0]
        screenspace = _screenspace_leaf(xyz)
        image, radii, posed_xyz, posed_cov, colors, opacity = _RenderFused.apply(
            xyz, log_scale, quat, opacity_logit, f_dc, f_rest, skin_wts, screenspace, bone_tf, campos,
            _settings(camera, bg_color, sh_degree, device), int(sh_degree), bool(isotropic), int(num_skinned), grad_sink, accumulate,
            bool(want_posed), plan)
        return {"render": torch.permute(image, (1, 2, 0)), "viewspace_points": screenspace, "visibility_filter": radii > 0, "radii": radii,
                "posed_xyz": posed_xyz, "posed_cov": posed_cov, "colors": colors, "cano_opacity": opacity}
    posed_xyz, posed_cov, colors, opacity = pose_gaussians(xyz, log_scale, quat, opacity_logit, f_dc, f_rest, skin_wts, bone_tf,
                                                          campos, sh_degree, isotropic, num_skinned, grad_sink, accumulate)
    out = render_gaussians(posed_xyz, posed_cov, xyz, None, opacity, camera, bg_color, colors_precomp=colors,
                           sh_degree=sh_degree, device=device, _cached_screenspace=True)
    out.update(posed_xyz=posed_xyz, posed_cov=posed_cov, colors=colors, cano_opacity=opacity)
    return out
