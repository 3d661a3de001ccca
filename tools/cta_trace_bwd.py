#!/usr/bin/env python
"""Timeline of the backward tile kernel's CTAs (work items = (tile, segment)) on the headline scene; needs a -DMB_TRACE_CTA
variant build:  MANUS_B200_LIB=manus_b200/lib/variants/libtrace.so python tools/cta_trace_bwd.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manus_b200 import _lib, synth  # noqa: E402
from manus_b200.dist import SceneRenderer  # noqa: E402

dev = torch.device("cuda", 0)
scene = synth.make_composite(500_000, seed=0)
r = SceneRenderer(scene, dev, 1920, 1080)
G = torch.rand(1080, 1920, 3, device=dev)
L = _lib.lib()
L.mb_debug_cta_trace.restype = C.c_int
L.mb_debug_cta_trace.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
for view in (0, 25, 40):
    for _ in range(3):
        out = r.render(view, sink=r.flat.grads)
        (out["render"] * G).sum().backward()
    torch.cuda.synchronize()
    n = 1 << 15
    t = np.zeros((n, 2), np.uint64)
    w = np.zeros(n, np.uint32)
    assert L.mb_debug_cta_trace(1, t.ctypes.data, w.ctypes.data, n) == 0
    ran = t[:, 1] > 0
    t, w = t[ran], w[ran]
    t0 = t[:, 0].min()
    st, en = (t[:, 0] - t0).astype(np.float64) / 1e3, (t[:, 1] - t0).astype(np.float64) / 1e3
    dur = en - st
    total = en.max()
    print(f"view {view}: {ran.sum()} items, kernel span {total:.1f} us; CTA durations p50/p90/p99/max = "
          f"{np.percentile(dur, 50):.1f}/{np.percentile(dur, 90):.1f}/{np.percentile(dur, 99):.1f}/{dur.max():.1f} us; "
          f"sum of CTA time {dur.sum() / 1e3:.2f} ms = {dur.sum() / total / 148:.1f} CTAs per SM on average")
    for lo, hi in ((1, 8), (9, 32), (33, 128), (129, 256), (257, 511), (512, 512)):
        m = (w >= lo) & (w <= hi)
        if m.any():
            print(f"   segment length {lo:3d}-{hi:3d}: {int(m.sum()):5d} items, mean {dur[m].mean():6.1f} us, CTA time {dur[m].sum() / 1e3:6.2f} ms "
                  f"({100 * dur[m].sum() / dur.sum():4.1f} %), entries {int(w[m].sum())}")
    for frac in (0.1, 0.3, 0.5, 0.7, 0.8, 0.9, 0.95):
        tt = total * frac
        print(f"   at {frac * 100:.0f}% of the span ({tt:.1f} us): {int(((st <= tt) & (en > tt)).sum())} CTAs running, {int((st > tt).sum())} not started")
