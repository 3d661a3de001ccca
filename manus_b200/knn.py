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


def nearest_point(queries: torch.Tensor, refs: torch.Tensor):
    """-> (dist [N] float32, index [N] int32): distance to and index of the nearest of ``refs`` [M,3] for every row of
    ``queries`` [N,3] (exact; ties -> lowest index)."""
    L = _lib.lib()
    if not queries.is_cuda:
        raise _lib.ManusB200Error("nearest_point needs CUDA tensors (there is no CPU path)")
    q, r = queries.detach().float().contiguous(), refs.detach().to(queries.device).float().contiguous()
    if q.dim() != 2 or q.shape[1] != 3 or r.dim() != 2 or r.shape[1] != 3:
        raise RuntimeError("pt1 and pt2 must have dimensions (num_points, 3)")
    if r.shape[0] == 0:
        raise RuntimeError("nearest_point: the reference set is empty")
    n, m = q.shape[0], r.shape[0]
    dist = torch.zeros(n, dtype=torch.float32, device=q.device)
    idx = torch.zeros(n, dtype=torch.int32, device=q.device)
    if n:
        with torch.cuda.device(q.device):
            ws = torch.empty(L.mb_nearest_workspace_bytes(n, m), dtype=torch.uint8, device=q.device)
            _lib.check(L.mb_nearest_point(ptr(q), n, ptr(r), m, ptr(dist), ptr(idx), ptr(ws), ws.numel(),
                                          torch.cuda.current_stream(q.device).cuda_stream), "mb_nearest_point")
    return dist, idx


def get_contact_dist(pt1: torch.Tensor, pt2: torch.Tensor):
    """/root/reference/src/utils/gaussian_utils.py:521-554: (contact_map [N] float32, contact_indices [N] float32 -- the
    reference stores the indices in a float32 taichi array) of the nearest point of pt2 for every point of pt1."""
    dist, idx = nearest_point(pt1, pt2)
    return dist, idx.float()


def get_contact_map(pt1: torch.Tensor, pt2: torch.Tensor, chunk: int = 1024) -> torch.Tensor:
    """/root/reference/src/utils/gaussian_utils.py:514-518 (chunked torch.cdist(...).min(1)); ``chunk`` is accepted and unused."""
    return nearest_point(pt1, pt2)[0]
