"""The gradient exchange of the view-sharded step over NVSwitch multicast memory (csrc/exchange.cu).

``MulticastExchange`` moves the flat gradient buffer of a ``FlatGaussians`` into a symmetric allocation (same offsets on every
rank, mapped through an NVSwitch multicast address; ``torch.distributed._symmetric_memory`` does the allocation and the handle
exchange -- plumbing, like the process group itself) and sums pieces of it over the ranks with this repository's own kernel:
rank r reduces 1/R of every piece with ``multimem.ld_reduce`` and broadcasts the sums with ``multimem.st``.  The result is the
same on every rank bit for bit.  There is no fallback inside this class: where multicast is unavailable the caller keeps the
NCCL all-reduce (``manus_b200.dist.PipelinedStep(exchange=None)``)."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .dist import PARAM_ORDER, FlatGaussians


class MulticastExchange:
    def __init__(self, flat: FlatGaussians, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        if not (dist.is_available() and dist.is_initialized()):
            raise _lib.ManusB200Error("MulticastExchange needs an initialised process group")
        self.group = group if group is not None else dist.group.WORLD
        dev = flat.grad.device
        with torch.cuda.device(dev):
            buf = symm_mem.empty(flat.grad.numel(), dtype=torch.float32, device=dev)
            self.hdl = symm_mem.rendezvous(buf, self.group)
        if not getattr(self.hdl, "multicast_ptr", 0):
            raise _lib.ManusB200Error("this allocation has no NVSwitch multicast address (multimem unavailable on this system)")
        buf.zero_()
        self.buf, self.flat = buf, flat
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        self.base = int(self.hdl.multicast_ptr) + int(getattr(self.hdl, "offset", 0))
        # the gradient views of `flat` now live in the symmetric buffer
        flat.grad = buf
        off = 0
        self.offsets = {}
        for name in PARAM_ORDER:
            cnt = flat.grads[name].numel()
            self.offsets[name] = off
            flat.grads[name] = buf[off: off + cnt].view(flat.grads[name].shape)
            off += cnt

    def pieces(self, lo: int, hi: int) -> List[Tuple[int, int]]:
        """(offset, count) in floats of the six per-parameter pieces that hold the gradients of Gaussians [lo, hi)."""
        out = []
        for name in PARAM_ORDER:
            w = self.flat.grads[name].numel() // self.flat.n
            out.append((self.offsets[name] + lo * w, (hi - lo) * w))
        return out

    def all_reduce(self, pieces: Sequence[Tuple[int, int]], max_ctas: int = 0, channel: int = 0) -> None:
        """SUM over the ranks of the given pieces, in place, enqueued on the current stream (two cross-GPU barriers around the
        kernel: every rank's gradients are complete before anybody reduces, every rank's sums have landed before anybody goes on)."""
        L = _lib.lib()
        n = len(pieces)
        offs = (C.c_int64 * n)(*[p[0] for p in pieces])
        cnts = (C.c_int64 * n)(*[p[1] for p in pieces])
        dev = self.buf.device
        with torch.cuda.device(dev):
            self.hdl.barrier(channel=channel)
            _lib.check(L.mb_multimem_allreduce(self.base, offs, cnts, n, self.rank, self.world, int(max_ctas),
                                               torch.cuda.current_stream(dev).cuda_stream), "mb_multimem_allreduce")
            self.hdl.barrier(channel=channel)

    def all_reduce_all(self, max_ctas: int = 0) -> None:
        self.all_reduce([(0, self.buf.numel())], max_ctas)
