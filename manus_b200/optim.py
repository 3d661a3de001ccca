"""Fused Adam on the flat parameter buffer.

Mirrors what ``GaussianModel.training_setup`` builds (/root/reference/src/models/gaussian.py:129-146): ONE
``torch.optim.Adam(l, lr=0.0, eps=1e-15)`` over six param groups that differ only in their learning rate
(xyz, f_dc, f_rest, opacity, scaling, rotation), with the xyz learning rate rescheduled every step
(``update_learning_rate``).  ``FlatAdam`` keeps both moments in flat buffers laid out like
``manus_b200.dist.FlatGaussians`` and performs the whole step in one kernel; ``step(shard=(begin, end))`` updates only a
slice, which is what the ZeRO-1 style data-parallel step (``sharded_adam_step``) uses:
reduce-scatter(grad) -> Adam on the rank's shard -> all-gather(param): the same bytes on the wire as the all-reduce,
1/R of the optimizer traffic per GPU.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import ptr
from .dist import PARAM_ORDER, FlatGaussians

# reference name of each flat segment's param group (gaussian.py:133-140) -> key in FlatGaussians
GROUP_OF = {"xyz": "xyz", "f_dc": "f_dc", "f_rest": "f_rest", "opacity": "opacity_logit", "scaling": "log_scale", "rotation": "quat"}


class FlatAdam:
    """Adam(betas=(0.9, 0.999), eps=1e-15, no weight decay, no amsgrad) over a FlatGaussians buffer, per-group learning rates."""

    def __init__(self, flat: FlatGaussians, lrs: Dict[str, float], betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-15):
        self.flat = flat
        self.betas, self.eps = betas, eps
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)
        self.step_count = 0
        self.lr = {}
        for ref_name, key in GROUP_OF.items():
            if ref_name not in lrs and key not in lrs:
                raise KeyError(f"learning rate for param group '{ref_name}' missing")
            self.lr[key] = float(lrs.get(ref_name, lrs.get(key)))
        ends, off = [], 0
        for name in PARAM_ORDER:
            off += flat.params[name].numel()
            ends.append(off)
        self._seg_end = (C.c_int64 * len(ends))(*ends)
        self.numel = off

    def rebind(self, flat: FlatGaussians, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor) -> None:
        """After densification / pruning changed N: new flat buffers and the moments that were carried over
        (manus_b200.densify.GaussianState); the step count and learning rates continue."""
        self.flat, self.exp_avg, self.exp_avg_sq = flat, exp_avg, exp_avg_sq
        ends, off = [], 0
        for name in PARAM_ORDER:
            off += flat.params[name].numel()
            ends.append(off)
        self._seg_end = (C.c_int64 * len(ends))(*ends)
        self.numel = off

    def set_lr(self, name: str, lr: float) -> None:
        """e.g. the per-step exponential schedule of the xyz group (gaussian.py:142-146, update_learning_rate)."""
        self.lr[GROUP_OF.get(name, name)] = float(lr)

    def step(self, shard: Optional[Tuple[int, int]] = None, grad: Optional[torch.Tensor] = None, grad_scale: float = 1.0,
             advance: bool = True) -> None:
        """One Adam step on elements [begin, end) (default: everything).  ``grad``: flat gradient buffer to read
        (default flat.grad; a reduce-scattered shard is passed as a full-size view by the caller).  All shards of one
        optimisation step must use the same step count: pass advance=False for the 2nd.. shard calls of a step."""
        L = _lib.lib()
        if not self.flat.data.is_cuda:
            raise _lib.ManusB200Error("FlatAdam needs CUDA buffers (there is no CPU path)")
        if advance:
            self.step_count += 1
        begin, end = (0, self.numel) if shard is None else shard
        g = self.flat.grad if grad is None else grad
        lrs = (C.c_double * len(PARAM_ORDER))(*[self.lr[k] for k in PARAM_ORDER])
        dev = self.flat.data.device
        with torch.cuda.device(dev):
            _lib.check(L.mb_fused_adam(ptr(self.flat.data), ptr(g), ptr(self.exp_avg), ptr(self.exp_avg_sq), int(begin), int(end),
                                       len(PARAM_ORDER), self._seg_end, lrs, self.step_count, self.betas[0], self.betas[1], self.eps,
                                       float(grad_scale), torch.cuda.current_stream(dev).cuda_stream), "mb_fused_adam")


def get_expon_lr_func(lr_init: float, lr_final: float, lr_delay_steps: int = 0, lr_delay_mult: float = 1.0,
                      max_steps: int = 1000000) -> Callable[[int], float]:
    """The xyz group's per-step learning rate (/root/reference/src/utils/gaussian_utils.py:212-247, used through
    ``xyz_scheduler_args`` at src/models/gaussian.py:143-146 and ``update_learning_rate`` :505-511): log-linear interpolation
    from lr_init (step 0) to lr_final (max_steps), optionally eased in over lr_delay_steps by
    lr_delay_mult + (1 - lr_delay_mult) sin(pi/2 clip(step / lr_delay_steps)); 0 for negative steps or when both ends are 0.
    Host arithmetic in float64 like the reference; feed the result to ``FlatAdam.set_lr("xyz", lr)`` before each step."""
    log_init, log_final = (math.log(lr_init), math.log(lr_final)) if lr_init > 0.0 and lr_final > 0.0 else (None, None)

    def lr_at(step: int) -> float:
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        delay = 1.0
        if lr_delay_steps > 0:
            delay = lr_delay_mult + (1.0 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
        t = min(max(step / max_steps, 0.0), 1.0)
        if log_init is None:      # one end is 0: the reference's log() gives -inf and exp() 0 (or NaN at the 0-weighted end)
            return delay * float(np.exp(np.log(np.float64(lr_init)) * (1 - t) + np.log(np.float64(lr_final)) * t))
        return delay * math.exp(log_init * (1.0 - t) + log_final * t)

    return lr_at


def shard_range(numel: int, rank: int, world_size: int, align: int = 4) -> Tuple[int, int]:
    """Contiguous slice of the flat buffer owned by ``rank`` (multiples of ``align`` elements; the last rank takes the rest)."""
    per = (numel + world_size - 1) // world_size
    per = (per + align - 1) // align * align
    return min(rank * per, numel), min((rank + 1) * per, numel)


def sharded_adam_step(opt: FlatAdam, n_views: int = 1, group=None) -> None:
    """Data-parallel optimiser step without replicated optimizer work: every rank holds local gradients in flat.grad;
    reduce-scatter them, update the owned slice of the parameters, all-gather the parameters.  Equivalent to
    all_reduce(grad) / n_views followed by a full Adam step on every rank."""
    flat = opt.flat
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        opt.step(grad_scale=1.0 / n_views)
        return
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = shard_range(opt.numel, 0, world)[1]
    padded = per * world
    if getattr(opt, "_pad_numel", None) != padded:
        opt._pad_numel = padded
        opt._grad_pad = torch.zeros(padded, dtype=torch.float32, device=flat.grad.device)
        opt._param_pad = torch.zeros(padded, dtype=torch.float32, device=flat.grad.device)
        opt._grad_shard = torch.zeros(per, dtype=torch.float32, device=flat.grad.device)
    opt._grad_pad[: opt.numel].copy_(flat.grad)
    dist.reduce_scatter_tensor(opt._grad_shard, opt._grad_pad, op=dist.ReduceOp.SUM, group=group)
    begin, end = shard_range(opt.numel, rank, world)
    # the kernel addresses gradients by flat offset: view the shard so that element `begin` of the flat layout is shard[0]
    opt._grad_pad[begin:begin + per].copy_(opt._grad_shard)
    opt.step(shard=(begin, end), grad=opt._grad_pad, grad_scale=1.0 / n_views)
    opt._param_pad[: opt.numel].copy_(flat.data)
    dist.all_gather_into_tensor(opt._param_pad, opt._param_pad[rank * per:(rank + 1) * per].clone(), group=group)
    flat.data.copy_(opt._param_pad[: opt.numel])
