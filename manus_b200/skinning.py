"""``skinning_weights_from_voxel_grid``: same signature and result as /root/reference/src/utils/gaussian_utils.py:167-196, the
lookup ``HandGaussianModel.get_skin_weights`` runs every training step (src/models/hand_gaussian.py:65-76).  One CUDA
kernel forward, one backward (manus_b200/csrc/skin.cu); gradients flow to ``xyz`` and, when it requires grad, to
``grid_weights``."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ptr


class _SkinWeights(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, grid_center, grid_scale, grid_weights):
        L = _lib.lib()
        if not xyz.is_cuda:
            raise _lib.ManusB200Error("manus_b200.skinning needs CUDA tensors (there is no CPU path)")
        dev = xyz.device
        x = xyz.detach().float().contiguous()
        g = grid_weights.detach().to(dev).float().contiguous()
        if g.dim() != 4 or x.dim() != 2 or x.shape[1] != 3:
            raise RuntimeError(f"expected xyz [N,3] and grid_weights [D,H,W,C], got {tuple(x.shape)} and {tuple(g.shape)}")
        c = torch.as_tensor(grid_center).detach().to(dev).float().reshape(-1).contiguous()
        s = torch.as_tensor(grid_scale).detach().to(dev).float().reshape(-1).contiguous()
        if c.numel() != 3 or s.numel() != 3:
            raise RuntimeError("grid_center and grid_scale must hold 3 values each")
        D, H, W, C = g.shape
        out = torch.empty((x.shape[0], C), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.mb_skin_weights_forward(ptr(x), x.shape[0], ptr(g), D, H, W, C, ptr(c), ptr(s), ptr(out),
                                                 torch.cuda.current_stream(dev).cuda_stream), "mb_skin_weights_forward")
        ctx.saved = (x, g, c, s)
        ctx.need_grid = grid_weights.requires_grad
        ctx.xyz_shape = xyz.shape
        return out

    @staticmethod
    def backward(ctx, g_out):
        L = _lib.lib()
        x, g, c, s = ctx.saved
        dev = x.device
        D, H, W, C = g.shape
        go = g_out.float().contiguous()
        g_xyz = torch.empty_like(x)
        g_grid = torch.zeros_like(g) if ctx.need_grid else None
        with torch.cuda.device(dev):
            _lib.check(L.mb_skin_weights_backward(ptr(x), x.shape[0], ptr(g), D, H, W, C, ptr(c), ptr(s), ptr(go), ptr(g_xyz), ptr(g_grid),
                                                  torch.cuda.current_stream(dev).cuda_stream), "mb_skin_weights_backward")
        return g_xyz.reshape(ctx.xyz_shape), None, None, g_grid


def skinning_weights_from_voxel_grid(xyz, grid_center, grid_scale, grid_weights):
    """gaussian_utils.py:167-196.  xyz [N,3]; grid_center [3]; grid_scale [1,3] (or [3]); grid_weights [D,H,W,C] -> [N,C]."""
    return _SkinWeights.apply(xyz, grid_center, grid_scale, grid_weights)
