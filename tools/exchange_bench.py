#!/usr/bin/env python
"""Timing of the pieces of the multi-GPU gradient exchange (run under torchrun with N >= 2 ranks)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 500_000


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


flat = torch.randn(N * 59, device=dev)
head = flat[: N * 11]
send = torch.randn(N * 3 + 340, device=dev)
recv = torch.empty(world, N * 3 + 340, device=dev)
res = {
    "all_reduce 118 MB (flat buffer)": timeit(lambda: dist.all_reduce(flat)),
    "all_reduce 22 MB (non-SH front)": timeit(lambda: dist.all_reduce(head)),
    "all_gather 6 MB per rank": timeit(lambda: dist.all_gather_into_tensor(recv.view(-1), send)),
    "both, back to back": timeit(lambda: (dist.all_gather_into_tensor(recv.view(-1), send), dist.all_reduce(head))),
}
if rank == 0:
    print({k: round(v, 1) for k, v in res.items()}, "us, world", world)
dist.destroy_process_group()
