#!/usr/bin/env python
"""Micro-benchmark of mb_radix_sort_pairs (CUDA events):  python tools/sort_bench.py [lib.so ...]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load(path):
    h = C.CDLL(path)
    h.mb_sort_workspace_bytes.restype = C.c_size_t
    h.mb_sort_workspace_bytes.argtypes = [C.c_int64]
    h.mb_radix_sort_pairs.restype = C.c_int
    h.mb_radix_sort_pairs.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
    return h


def bench(h, keys, end_bit, dev_count, reps=20):
    n = keys.numel()
    cap = int(n * 1.1) if dev_count else n
    kin = torch.zeros(cap, dtype=torch.int32, device="cuda"); kin[:n] = keys
    vin = torch.arange(cap, dtype=torch.int32, device="cuda")
    ko, vo = torch.empty_like(kin), torch.empty_like(vin)
    ws = torch.empty(h.mb_sort_workspace_bytes(cap), dtype=torch.uint8, device="cuda")
    ndev = torch.tensor([n], dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")

    def run():
        if dev_count:
            return h.mb_radix_sort_pairs(kin.data_ptr(), vin.data_ptr(), ko.data_ptr(), vo.data_ptr(), -1, ndev.data_ptr(), cap, end_bit, ws.data_ptr(), ws.numel(), s)
        return h.mb_radix_sort_pairs(kin.data_ptr(), vin.data_ptr(), ko.data_ptr(), vo.data_ptr(), n, None, n, end_bit, ws.data_ptr(), ws.numel(), s)

    for _ in range(3):
        assert run() == 0
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ok = bool((ko[:n].cpu().numpy().view(np.uint32) == np.sort(keys.cpu().numpy().view(np.uint32), kind="stable")).all())
    return float(np.median(ts)) * 1e3, ok


def main():
    libs = sys.argv[1:] or [os.path.join(ROOT, "manus_b200", "lib", "libmanus_b200.so")]
    rng = np.random.default_rng(0)
    depth = rng.uniform(1.1, 1.6, 500_000).astype(np.float32).view(np.int32)
    tiles = rng.integers(0, 8160, 2_300_000).astype(np.int32)
    kd, kt = torch.tensor(depth, device="cuda"), torch.tensor(tiles, device="cuda")
    for path in libs:
        h = load(path)
        a, ok1 = bench(h, kd, 32, False)
        b, ok2 = bench(h, kt, 13, True)
        print(f"{os.path.basename(path):40s} depth 500k x 32 bit: {a:8.1f} us ({'ok' if ok1 else 'WRONG'})   tiles 2.3M x 13 bit: {b:8.1f} us ({'ok' if ok2 else 'WRONG'})", flush=True)


if __name__ == "__main__":
    main()
