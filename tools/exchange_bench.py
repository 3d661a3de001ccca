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
# ---- the repository's own exchange over NVSwitch multicast memory against NCCL on the same 118 MB
from manus_b200.dist import FlatGaussians  # noqa: E402
from manus_b200.exchange import MulticastExchange  # noqa: E402

fg = FlatGaussians(N, dev)
try:
    ex = MulticastExchange(fg)
    src = torch.randn(N * 59, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))
    fg.grad.copy_(src)
    want = src.clone()
    dist.all_reduce(want)
    torch.cuda.synchronize(); dist.barrier()
    ex.all_reduce_all()
    torch.cuda.synchronize()
    err = float((fg.grad - want).abs().max()) / float(want.abs().max())
    same = torch.equal(fg.grad, want)
    # every rank must hold the same bits
    probe = fg.grad[:: 4099].clone()
    ref = probe.clone()
    dist.broadcast(ref, 0)
    identical = bool(torch.equal(probe, ref))
    for ctas in (0, 64, 32, 16):
        res[f"multimem all-reduce 118 MB, max_ctas={ctas or 'SMs'}"] = timeit(lambda: ex.all_reduce_all(ctas))
    # hybrid: part of every piece through plain peer-to-peer loads / stores beside the in-switch reduction
    whole = [(0, N * 59)]
    for frac in (0.3, 0.4, 0.5, 0.6):
        fg.grad.copy_(src)
        torch.cuda.synchronize(); dist.barrier()
        ex.all_reduce_hybrid(whole, frac, 32, 64)
        torch.cuda.synchronize()
        herr = float((fg.grad - want).abs().max()) / float(want.abs().max())
        for pc in (48, 128):
            res[f"hybrid p2p={frac} p2p_ctas={pc}"] = timeit(lambda: ex.all_reduce_hybrid(whole, frac, 32, pc))
        res[f"hybrid p2p={frac} max rel err"] = herr
    res["p2p only"] = timeit(lambda: ex.all_reduce_hybrid(whole, 1.0, 32, 128))
    chunk = ex.pieces(0, 125056)
    for frac in (0.4, 0.5):
        res[f"hybrid p2p={frac}, one of 4 ranges"] = timeit(lambda: ex.all_reduce_hybrid(chunk, frac, 32, 64))
    res["multimem all-reduce, one of 4 ranges (6 pieces, 29.5 MB)"] = timeit(lambda: ex.all_reduce(chunk))
    res["max rel err vs NCCL"] = err
    res["bitwise equal to NCCL"] = float(same)
    res["replicas bitwise identical"] = float(identical)
except Exception as e:  # noqa: BLE001
    res["multimem"] = f"unavailable: {type(e).__name__}: {e}"
if rank == 0:
    print({k: (round(v, 1) if isinstance(v, float) and v > 1 else v) for k, v in res.items()}, "us, world", world)
dist.destroy_process_group()
