"""Fused pre-raster step (LBS + covariance + SH->RGB + activations) with a full backward.

The reference has no operator boundary here -- it is inline PyTorch:
  LBS of means / covariances     /root/reference/src/modules/hand_dynamic.py:86-137
  covariance build, activations  /root/reference/src/models/gaussian.py:48-93
  SH -> RGB (canonical view dir) /root/reference/src/utils/gaussian_utils.py:431-449, src/utils/sh_utils.py:57-120
``pose_gaussians`` takes the raw parameters of ``GaussianModel`` (same names as its nn.Parameters) plus the per-frame
skin weights / bone transforms / camera centre and returns exactly what ``render_gaussians`` feeds the rasterizer.
One CUDA kernel forward, one backward (manus_b200/csrc/pose.cu).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import ptr


def bone_transforms(bones_posed: torch.Tensor, bones_rest: torch.Tensor, append_identity: bool = True,
                    rest_inv: Optional[torch.Tensor] = None) -> torch.Tensor:
    """T_b = posed_b . inv(rest_b) (+ identity background bone): hand_dynamic.py:93-102.  20 tiny matrices: plumbing.
    The rest pose is constant for a scene, so callers that render many frames pass ``rest_inv = torch.linalg.inv(bones_rest)``
    computed once (same values; saves the batched LU kernels of every frame)."""
    if rest_inv is None:
        rest_inv = torch.linalg.inv(bones_rest)
    tfs = torch.einsum("nij,njk->nik", bones_posed, rest_inv)
    if append_identity:
        tfs = torch.cat([tfs, torch.eye(4, dtype=tfs.dtype, device=tfs.device)[None]], dim=0)
    return tfs


def _f32c(t):
    return None if t is None else t.detach().float().contiguous()


def _inputs(xyz, log_scale, quat, opacity_logit, f_dc, f_rest, skin_wts, bone_tf, campos, sh_degree, isotropic, num_skinned):
    pi = _lib.PoseInputs()
    N = xyz.shape[0]
    pi.num_points = N
    pi.num_skinned = num_skinned
    pi.num_bones = 0 if skin_wts is None else skin_wts.shape[1]
    pi.sh_degree = int(sh_degree)
    pi.sh_coeffs = 1 + (0 if f_rest is None else f_rest.shape[1])
    pi.isotropic = int(bool(isotropic))
    pi.xyz, pi.log_scale, pi.quat, pi.opacity_logit = ptr(xyz), ptr(log_scale), ptr(quat), ptr(opacity_logit)
    pi.f_dc, pi.f_rest = ptr(f_dc), (ptr(f_rest) if f_rest is not None and f_rest.numel() else None)
    pi.skin_wts, pi.campos = ptr(skin_wts), ptr(campos)
    if isinstance(bone_tf, (tuple, list)):      # (bones_posed, bones_rest_inv): the kernels build T_b = posed_b rest_b^-1 (+ identity rows)
        posed, rest_inv = bone_tf[0], bone_tf[1]
        pi.bone_tf = None
        pi.bones_posed, pi.bones_rest_inv, pi.num_posed_bones = ptr(posed), ptr(rest_inv), int(rest_inv.shape[0])
    else:
        pi.bone_tf = ptr(bone_tf)
    return pi


class _PoseGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, log_scale, quat, opacity_logit, f_dc, f_rest, skin_wts, bone_tf, campos, sh_degree, isotropic,
                num_skinned, grad_sink, accumulate=False):
        L = _lib.lib()
        if not xyz.is_cuda:
            raise _lib.ManusB200Error("manus_b200.pose_gaussians needs CUDA tensors (there is no CPU path)")
        dev = xyz.device
        t = [_f32c(v) for v in (xyz, log_scale, quat, opacity_logit, f_dc, f_rest, skin_wts, bone_tf)]
        cam = _f32c(campos).reshape(-1)[:3].contiguous()
        N = t[0].shape[0]
        if t[6] is not None:
            if t[6].shape[0] != num_skinned or t[7] is None or t[7].shape[0] != t[6].shape[1]:
                raise RuntimeError(f"skin_wts {tuple(t[6].shape)} does not match num_skinned={num_skinned} / bone_tf "
                                   f"{None if t[7] is None else tuple(t[7].shape)}")   # hand_dynamic.py:104
        pi = _inputs(*t, cam, sh_degree, isotropic, num_skinned)
        new = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        posed_xyz, cov6, colors, opacity = new(N, 3), new(N, 6), new(N, 3), new(N, 1)
        with torch.cuda.device(dev):
            _lib.check(L.mb_pose_forward(C.byref(pi), ptr(posed_xyz), ptr(cov6), ptr(colors), ptr(opacity), None,
                                         torch.cuda.current_stream(dev).cuda_stream), "mb_pose_forward")
        ctx.saved = (t, cam, sh_degree, isotropic, num_skinned)
        ctx.grad_sink = grad_sink
        ctx.accumulate = bool(accumulate) and grad_sink is not None
        ctx.need_skin = skin_wts is not None and skin_wts.requires_grad
        ctx.shapes = [None if v is None else v.shape for v in (xyz, log_scale, quat, opacity_logit, f_dc, f_rest, skin_wts)]
        return posed_xyz, cov6, colors, opacity

    @staticmethod
    def backward(ctx, g_xyz_p, g_cov6, g_colors, g_opacity):
        L = _lib.lib()
        t, cam, sh_degree, isotropic, num_skinned = ctx.saved
        dev = t[0].device
        N = t[0].shape[0]
        pi = _inputs(*t, cam, sh_degree, isotropic, num_skinned)
        zeros = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        gin = [_f32c(g) if g is not None else zeros(*s) for g, s in
               ((g_xyz_p, (N, 3)), (g_cov6, (N, 6)), (g_colors, (N, 3)), (g_opacity, (N, 1)))]
        sink = ctx.grad_sink
        if sink is not None:
            # write straight into caller-owned dense buffers (e.g. the flat all-reduce buffer): no autograd accumulation pass
            # sink["f_rest"] may be None: the per-view SH gradient is rank one and can be rebuilt from g_f_dc (sh_grad_from_views)
            g_xyz, g_ls, g_q, g_ol, g_fdc, g_fr = (sink.get(k) for k in ("xyz", "log_scale", "quat", "opacity_logit", "f_dc", "f_rest"))
            for g, ref in ((g_xyz, t[0]), (g_ls, t[1]), (g_q, t[2]), (g_ol, t[3]), (g_fdc, t[4]), (g_fr, t[5])):
                if g is None and ref is t[5]:
                    continue
                if ref is not None and (g is None or g.numel() != ref.numel() or not g.is_contiguous() or g.dtype != torch.float32):
                    raise RuntimeError("grad_sink tensors must be dense fp32 with the parameter's size")
        else:
            new = lambda ref: torch.empty_like(ref)
            g_xyz, g_ls, g_q, g_ol, g_fdc = new(t[0]), new(t[1]), new(t[2]), new(t[3]), new(t[4])
            g_fr = new(t[5]) if t[5] is not None else None
        g_skin = torch.empty_like(t[6]) if ctx.need_skin else None
        with torch.cuda.device(dev):
            if sink is not None and sink.get("_wait") is not None:      # ordered accumulation: after the previous view's launch
                torch.cuda.current_stream(dev).wait_event(sink["_wait"])
            entry = L.mb_pose_backward_accumulate if ctx.accumulate else L.mb_pose_backward
            if ctx.accumulate and g_skin is not None:
                g_skin.zero_()
            _lib.check(entry(C.byref(pi), ptr(gin[0]), ptr(gin[1]), ptr(gin[2]), ptr(gin[3]), ptr(g_xyz), ptr(g_ls),
                             ptr(g_q), ptr(g_ol), ptr(g_fdc), ptr(g_fr) if g_fr is not None and g_fr.numel() else None,
                             ptr(g_skin), torch.cuda.current_stream(dev).cuda_stream), "mb_pose_backward")
            if sink is not None and sink.get("_record") is not None:
                sink["_record"].record(torch.cuda.current_stream(dev))
        sh = ctx.shapes
        rs = lambda g, s: None if g is None else g.reshape(s)
        if sink is not None:
            return (None,) * 6 + (rs(g_skin, sh[6]), None, None, None, None, None, None, None)
        return (rs(g_xyz, sh[0]), rs(g_ls, sh[1]), rs(g_q, sh[2]), rs(g_ol, sh[3]), rs(g_fdc, sh[4]), rs(g_fr, sh[5]),
                rs(g_skin, sh[6]), None, None, None, None, None, None, None)


def pose_gaussians(xyz, log_scale, quat, opacity_logit, f_dc, f_rest, skin_wts: Optional[torch.Tensor],
                   bone_tf: Optional[torch.Tensor], campos, sh_degree: int = 3, isotropic: bool = False,
                   num_skinned: Optional[int] = None, grad_sink: Optional[dict] = None, accumulate: bool = False):
    """-> (posed_xyz[N,3], posed_cov6[N,6], colors[N,3], opacity[N,1]).

    xyz [N,3], log_scale [N,3] ([N,1] if isotropic), quat [N,4] raw, opacity_logit [N,1], f_dc [N,1,3],
    f_rest [N,K-1,3]: the nn.Parameters of GaussianModel (gaussian.py:120-125).
    skin_wts [n,B] + bone_tf [B,4,4] (``bone_transforms``): Gaussians [0,n) are skinned, the rest static (tf = I);
    both None -> object path (src/modules/object.py:32-41).  campos [3] or [1,3].
    grad_sink: optional {"xyz","log_scale","quat","opacity_logit","f_dc","f_rest"} -> dense fp32 tensors; when given, the
    backward kernel OVERWRITES them with the parameter gradients and autograd receives no gradient for those inputs
    (used by the data-parallel step: the sinks are views of the flat all-reduce buffer).
    accumulate (with grad_sink): the backward kernel ADDS to the sinks instead (gradient accumulation over the views of one
    step, hand_dynamic.py:248,259-277; bulk TMA reduce-add, no read-modify-write pass).  Views running on different streams
    order their accumulating launches through two optional entries of grad_sink: "_wait" (a torch.cuda.Event the backward
    launch waits for) and "_record" (an event recorded right after it).
    """
    if num_skinned is None:
        num_skinned = 0 if skin_wts is None else skin_wts.shape[0]
    return _PoseGaussians.apply(xyz, log_scale, quat, opacity_logit, f_dc, f_rest, skin_wts, bone_tf, campos, int(sh_degree),
                                bool(isotropic), int(num_skinned), grad_sink, bool(accumulate))


def sh_grad_from_views(xyz, skin_wts, num_skinned, sh_degree, sh_coeffs, bone_tf_all, campos_all, g_f_dc_all, out_f_dc, out_f_rest):
    """SH-coefficient gradients of a sum over R views from the views' DC gradients (see mb_sh_grad_from_views):
    bone_tf_all [R,B,4,4] (or None for static scenes), campos_all [R,3], g_f_dc_all [R,N,3] -> out_f_dc [N,1,3], out_f_rest [N,K-1,3].
    The three per-view inputs may be views into one [R, stride] buffer (one gathered record per rank)."""
    L = _lib.lib()
    dev = xyz.device
    R = g_f_dc_all.shape[0]
    x = _f32c(xyz)
    sk = _f32c(skin_wts) if num_skinned else None
    B = 0 if sk is None else sk.shape[1]
    stride = 0
    if R > 1 and g_f_dc_all.stride(0) != g_f_dc_all[0].numel():
        stride = g_f_dc_all.stride(0)          # strided record layout: every per-view array must share the stride
        if campos_all.stride(0) != stride or (num_skinned and bone_tf_all.stride(0) != stride):
            raise RuntimeError("strided per-view inputs must be views of one [R, stride] buffer")
        bt, cp, gd = (bone_tf_all if num_skinned else None), campos_all, g_f_dc_all
    else:
        bt = _f32c(bone_tf_all) if num_skinned else None
        cp, gd = _f32c(campos_all), _f32c(g_f_dc_all)
    with torch.cuda.device(dev):
        _lib.check(L.mb_sh_grad_from_views(ptr(x), x.shape[0], ptr(sk), int(num_skinned), B, int(sh_degree), int(sh_coeffs), R, ptr(bt),
                                           ptr(cp), ptr(gd), int(stride), ptr(out_f_dc), ptr(out_f_rest) if out_f_rest is not None and out_f_rest.numel() else None,
                                           torch.cuda.current_stream(dev).cuda_stream), "mb_sh_grad_from_views")
