#!/usr/bin/env python
"""How much do independent views gain from running concurrently?  V frames (forward + backward, own renderer state and
gradient sink each) captured on V streams of ONE CUDA graph; prints ms per frame for V = 1, 2, 3, 4.
python tools/overlap_probe.py [gaussians]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manus_b200 import rasterizer as rz, synth  # noqa: E402
from manus_b200.dist import SceneRenderer  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
W, H = 1920, 1080
dev = torch.device("cuda", 0)
scene = synth.make_composite(N, seed=0)
VMAX = 4
rs = [SceneRenderer(scene, dev, W, H) for _ in range(VMAX)]
G = torch.rand(H, W, 3, device=dev)
staged = {}
for v in range(8):
    _, c, b = rs[0].view_inputs_host(v)
    staged[v] = (c.to(dev), b.to(dev))
rz.set_capacity_mode("exact")
dmax = 0
for v in range(8):
    out = rs[0].render(v, cam_dev=staged[v][0], bones_dev=staged[v][1])
    dmax = max(dmax, rz.check_overflow())
rz.set_capacity_mode("reserve", margin=1.1)
rz.reserve_capacity(0, scene.n, H, W, dmax)


def frame(i, slot):
    r = rs[i]
    out = r.render(0, sink=r.flat.grads, cam_dev=slot[0], bones_dev=slot[1], device_intrinsics=True)
    (out["render"] * G).sum().backward()


for V in range(1, VMAX + 1):
    slots = [(staged[i][0].clone(), staged[i][1].clone()) for i in range(V)]
    sides = [torch.cuda.Stream(device=dev) for _ in range(V)]
    for i in range(V):
        with torch.cuda.stream(sides[i]):
            for _ in range(2):
                frame(i, slots[i])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream(dev)
        for s in sides[1:]:
            s.wait_stream(cur)
        frame(0, slots[0])
        for i in range(1, V):
            with torch.cuda.stream(sides[i]):
                frame(i, slots[i])
        for s in sides[1:]:
            cur.wait_stream(s)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 100
    e0.record()
    for it in range(K):
        for i in range(V):
            v = (it * V + i) % 8
            slots[i][0].copy_(staged[v][0], non_blocking=True)
            slots[i][1].copy_(staged[v][1], non_blocking=True)
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"V={V}: {ms:.4f} ms per graph, {ms / V:.4f} ms per frame, {1e3 * V / ms:.0f} frames/s", flush=True)
    del g
