#!/usr/bin/env python
"""HBM bandwidth a plain kernel reaches on this GPU for read-only / write-only / copy streams (context for the roofline
fractions of the HBM-bound kernels).  python tools/bw_probe.py [build]"""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "probe", "libbw_probe.so")


def build():
    subprocess.check_call(["nvcc", "-O3", "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
                           os.path.join(HERE, "probe", "bw_probe.cu"), "-o", LIB])


def main():
    if not os.path.exists(LIB) or "build" in sys.argv:
        build()
    if "build" in sys.argv:
        return
    import torch
    lib = C.CDLL(LIB)
    lib.bw_probe.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_size_t, C.c_void_p]
    nbytes = 1 << 30
    a = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda").normal_()
    b = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    sms = torch.cuda.get_device_properties(0).multi_processor_count

    def run(mode, grid, chunk=0, total=nbytes, moved=None):
        for _ in range(3):
            lib.bw_probe(mode, a.data_ptr(), b.data_ptr(), total, grid, chunk, s)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); lib.bw_probe(mode, a.data_ptr(), b.data_ptr(), total, grid, chunk, s); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return (moved or total) / best / 1e6

    for name, mode, mult in (("read-only", 0, 1), ("write-only", 1, 1), ("copy (read+write bytes)", 2, 2)):
        print(name, {f"{k} CTAs/SM": round(run(mode, sms * k, moved=nbytes * mult)) for k in (4, 8, 16, 32)}, "GB/s")
    for chunk in (23040, 65536, 1 << 20):
        print(f"read-only, one {chunk}-B chunk per CTA turn", {f"{k} CTAs/SM": round(run(3, sms * k, chunk)) for k in (4, 8, 16)}, "GB/s")
    # a pose-kernel sized read (143 MB), cold
    print("read-only 143 MB", round(run(0, sms * 16, total=143 << 20)), "GB/s")


if __name__ == "__main__":
    main()
