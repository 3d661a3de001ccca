#!/usr/bin/env python
"""Timeline of the forward tile kernel's CTAs on the headline scene (needs a -DMB_TRACE_CTA variant build):
MANUS_B200_LIB=manus_b200/lib/variants/libtrace.so python tools/cta_trace.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manus_b200 import _lib, rasterizer as rz, synth  # noqa: E402
from manus_b200.dist import SceneRenderer  # noqa: E402
from manus_b200.pose import pose_gaussians  # noqa: E402

dev = torch.device("cuda", 0)
scene = synth.make_composite(500_000, seed=0)
r = SceneRenderer(scene, dev, 1920, 1080)
L = _lib.lib()
L.mb_debug_cta_trace.restype = C.c_int
L.mb_debug_cta_trace.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
for view in (0, 25, 40):
    cam, c, b = r.view_inputs_host(view)
    c, b = c.to(dev), b.to(dev)
    with torch.no_grad():
        bone_tf = torch.cat([torch.bmm(b.view(-1, 4, 4), r.rest_inv), r._eye], 0)
        px, pc, col, op = pose_gaussians(*[p.detach() for p in r.flat.leaves()], r.skin, bone_tf, c[32:35], 3, False, r.n_hand)
        s = rz.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, r.bg, 1.0, c[0:16], c[16:32], 3, c[32:35], False, False)
        for _ in range(3):
            rz.rasterize_forward(s, px, op.reshape(-1), colors_precomp=col, cov3D_precomp=pc)
    n = 16320
    t = np.zeros((n, 2), np.uint64)
    w = np.zeros(n, np.uint32)
    assert L.mb_debug_cta_trace(0, t.ctypes.data, w.ctypes.data, n) == 0
    t0 = t[:, 0].min()
    st, en = (t[:, 0] - t0).astype(np.float64) / 1e3, (t[:, 1] - t0).astype(np.float64) / 1e3
    dur = en - st
    total = en.max()
    # how much CTA-time is in flight as a function of time
    print(f"view {view}: kernel span {total:.1f} us; CTA durations p50/p99/max = {np.percentile(dur, 50):.1f}/{np.percentile(dur, 99):.1f}/{dur.max():.1f} us")
    for frac in (0.5, 0.7, 0.8, 0.9, 0.95):
        tt = total * frac
        running = int(((st <= tt) & (en > tt)).sum())
        print(f"   at {frac * 100:.0f}% of the span ({tt:.1f} us): {running} CTAs running, {int((st > tt).sum())} not started")
    top = np.argsort(-dur)[:8]
    print("   longest CTAs: " + ", ".join(f"#{i} start {st[i]:.1f} dur {dur[i]:.1f} consumed {w[i]}" for i in top))
