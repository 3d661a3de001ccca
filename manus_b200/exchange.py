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

    def all_reduce_hybrid(self, pieces: Sequence[Tuple[int, int]], p2p_fraction: float, max_ctas: int = 0, p2p_ctas: int = 0,
                          channel: int = 0) -> None:
        """The same sum with every piece split in two: the front goes through the switch (multimem kernel), the last
        ``p2p_fraction`` of it through plain peer-to-peer loads / stores (mb_p2p_allreduce) on a second stream, both inside the
        same pair of barriers.  The in-switch reduction leaves about half of the NVLink bandwidth idle; the peer-to-peer kernel
        uses it."""
        L = _lib.lib()
        mm, pp = [], []
        for off, cnt in pieces:
            tail = (int(cnt * p2p_fraction) // 4) * 4
            head = cnt - tail
            # the peer-to-peer part must start on a 16-byte boundary
            shift = (-(off + head)) % 4
            head, tail = head + shift, tail - shift
            if head > 0:
                mm.append((off, head))
            if tail >= 4 and tail % 4 == 0:
                pp.append((off + head, tail))
            elif tail > 0:
                mm[-1] = (off, cnt)
        dev = self.buf.device
        if not hasattr(self, "_side"):
            self._side = torch.cuda.Stream(device=dev)
            self._fork, self._join = torch.cuda.Event(), torch.cuda.Event()
            ptrs = [int(p) for p in self.hdl.buffer_ptrs]
            self._peers = (C.c_void_p * len(ptrs))(*ptrs)
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            self.hdl.barrier(channel=channel)
            self._fork.record(cur)
            if pp:
                with torch.cuda.stream(self._side):
                    self._side.wait_event(self._fork)
                    offs = (C.c_int64 * len(pp))(*[p[0] for p in pp])
                    cnts = (C.c_int64 * len(pp))(*[p[1] for p in pp])
                    _lib.check(L.mb_p2p_allreduce(self._peers, offs, cnts, len(pp), self.rank, self.world, int(p2p_ctas),
                                                  self._side.cuda_stream), "mb_p2p_allreduce")
                    self._join.record(self._side)
            if mm:
                offs = (C.c_int64 * len(mm))(*[p[0] for p in mm])
                cnts = (C.c_int64 * len(mm))(*[p[1] for p in mm])
                _lib.check(L.mb_multimem_allreduce(self.base, offs, cnts, len(mm), self.rank, self.world, int(max_ctas), cur.cuda_stream),
                           "mb_multimem_allreduce")
            if pp:
                cur.wait_event(self._join)
            self.hdl.barrier(channel=channel)

    def all_reduce_all(self, max_ctas: int = 0, p2p: bool = False) -> None:
        """The whole buffer.  p2p=True: through peer-to-peer loads / stores only (mb_p2p_allreduce) -- the faster way on TWO GPUs
        (238 us for 118 MB against 340 us through the switch's reduction and 252 us for NCCL); from four ranks on the in-switch
        reduction wins (302 us at eight ranks against 374 us peer-to-peer; mixing the two does not help: 322-356 us, the links
        are shared)."""
        if p2p:
            self.all_reduce_hybrid([(0, self.buf.numel())], 1.0, 0, 128)      # nothing runs beside it: enough CTAs to cover the link latency
        else:
            self.all_reduce([(0, self.buf.numel())], max_ctas)
