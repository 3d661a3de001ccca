"""``distCUDA2``: mean squared distance to the 3 nearest other points.

Drop-in for ``from simple_knn._C import distCUDA2`` (/root/reference/src/models/gaussian.py:4, called once at :110 to
initialise the log-scales).  points [N,3] float32 CUDA -> float32 [N].
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ptr


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    L = _lib.lib()
    if not points.is_cuda:
        raise _lib.ManusB200Error("distCUDA2 needs a CUDA tensor (there is no CPU path)")
    p = points.detach().float().contiguous()
    if p.dim() != 2 or p.shape[1] != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    n = p.shape[0]
    out = torch.zeros(n, dtype=torch.float32, device=p.device)
    if n == 0:
        return out
    with torch.cuda.device(p.device):
        ws = torch.empty(L.mb_knn_workspace_bytes(n), dtype=torch.uint8, device=p.device)
        _lib.check(L.mb_dist2_knn3(ptr(p), n, ptr(out), ptr(ws), ws.numel(), torch.cuda.current_stream(p.device).cuda_stream),
                   "mb_dist2_knn3")
    return out
