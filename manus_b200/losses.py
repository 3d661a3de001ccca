"""Image losses of the training step: same names and call pattern as /root/reference/src/utils/loss_utils.py
(``l1_loss``, ``ssim``) plus the fused ``photometric_loss`` that evaluates what ``loss_func``
(/root/reference/src/modules/base.py:323-365) builds from them -- ``w_l1 * l1 + w_ssim * (1 - ssim)`` -- with value and
gradient in one CUDA kernel (manus_b200/csrc/loss.cu).

The reference feeds HWC tensors (pred [H,W,3], gt [1,H,W,3]) to an SSIM written for CHW: ``channel = img.size(-3)`` is
the image height, so the 11x11 window filters every row over its (W, 3) plane.  These functions reproduce that behaviour
(it is what the trained models were optimised with); they are not a conventional SSIM.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ptr


def _hw3(t: torch.Tensor, what: str, dense: bool) -> torch.Tensor:
    if t.dim() == 4 and t.shape[0] == 1:
        t = t[0]
    if t.dim() != 3 or t.shape[-1] != 3:
        raise RuntimeError(f"{what} must be [H,W,3] (or [1,H,W,3]) like pred['render'] / batch['rgb'][..., :3], got {tuple(t.shape)}")
    if not t.is_cuda:
        raise _lib.ManusB200Error("manus_b200.losses needs CUDA tensors (there is no CPU path)")
    t = t.float()
    return t.contiguous() if dense else t      # pred is read with its strides (permuted [3,H,W] rasterizer output)


class _PhotometricLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, w_l1, w_ssim):
        L = _lib.lib()
        p, g = _hw3(pred.detach(), "pred", False), _hw3(gt.detach(), "gt", True)
        if p.shape != g.shape:
            raise RuntimeError(f"pred {tuple(p.shape)} and gt {tuple(g.shape)} differ")
        H, W, _ = p.shape
        dev = p.device
        out = torch.empty(3, dtype=torch.float32, device=dev)
        d_pred = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
        sy, sx, sc = p.stride()
        with torch.cuda.device(dev):
            ws = torch.empty(L.mb_photometric_loss_workspace_bytes(H, W), dtype=torch.uint8, device=dev)
            _lib.check(L.mb_photometric_loss(ptr(p), sy, sx, sc, ptr(g), H, W, float(w_l1), float(w_ssim), ptr(out), ptr(d_pred), ptr(ws), ws.numel(),
                                             torch.cuda.current_stream(dev).cuda_stream), "mb_photometric_loss")
        ctx.d_pred = d_pred
        ctx.pred_shape = pred.shape
        loss, l1, ss = out[0], out[1], out[2]
        ctx.mark_non_differentiable(l1, ss)
        return loss, l1, ss

    @staticmethod
    def backward(ctx, g_loss, _g_l1, _g_ss):
        d = ctx.d_pred
        ctx.d_pred = None
        return d.mul_(g_loss).reshape(ctx.pred_shape), None, None, None


def photometric_loss(pred: torch.Tensor, gt: torch.Tensor, w_l1: float = 0.8, w_ssim: float = 0.2, return_terms: bool = False):
    """``w_l1 * l1_loss(pred, gt) + w_ssim * (1 - ssim(pred, gt))`` (base.py:329-347 with the weights of config/*.yaml:22-23).
    pred [H,W,3] (requires grad), gt [H,W,3] or [1,H,W,3].  return_terms: also (mean |pred - gt|, mean ssim) for logging."""
    loss, l1, ss = _PhotometricLoss.apply(pred, gt, float(w_l1), float(w_ssim))
    return (loss, l1, ss) if return_terms else loss


def l1_loss(network_output: torch.Tensor, gt: torch.Tensor, mean: bool = True) -> torch.Tensor:
    """loss_utils.py:22-27.  The per-pixel map (mean=False) is plain elementwise work and stays in PyTorch."""
    if not mean:
        return torch.abs(network_output - gt)
    return _PhotometricLoss.apply(network_output, gt, 1.0, 0.0)[0]


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True) -> torch.Tensor:
    """loss_utils.py:57-97 for the reference's call ``ssim(pred[H,W,3], gt[1,H,W,3])`` (base.py:347)."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("manus_b200.losses.ssim implements the reference's only call: window_size=11, size_average=True")
    # loss = -1 * (1 - ssim) = ssim - 1
    return _PhotometricLoss.apply(img1, img2, 0.0, -1.0)[0] + 1.0
