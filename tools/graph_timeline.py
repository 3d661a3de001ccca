#!/usr/bin/env python
"""Timeline of the kernels of one replayed step (GraphedStep with V views in flight, headline scene): when does each
kernel of each view run, how long does it take under contention, how much of the step has k kernels in flight?
python tools/graph_timeline.py [V]"""
import os
import sys
from collections import defaultdict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manus_b200 import _lib, rasterizer as rz, synth  # noqa: E402
from manus_b200.dist import GraphedStep, SceneRenderer  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 4
W, H, NV = 1920, 1080, 50
dev = torch.device("cuda", 0)
scene = synth.make_composite(500_000, seed=0)
r = SceneRenderer(scene, dev, W, H)
G = torch.rand(H, W, 3, device=dev)
staged = {}
rz.set_capacity_mode("exact")
dmax = 0
for v in range(NV):
    _, c, b = r.view_inputs_host(v)
    staged[v] = (c.to(dev), b.to(dev))
    if v % 5 == 0:
        r.render(v, cam_dev=staged[v][0], bones_dev=staged[v][1])
        dmax = max(dmax, rz.check_overflow())
rz.set_capacity_mode("reserve", margin=1.4)
rz.reserve_capacity(0, scene.n, H, W, dmax)
def loss(image, target):      # the bench's probe loss: one dot product, its gradient (the target itself) seeds the backward
    return torch.dot(image.permute(2, 0, 1).reshape(-1), target.permute(2, 0, 1).reshape(-1)), target


G = G.permute(2, 0, 1).contiguous().permute(1, 2, 0)       # stored CHW like the rendered image
step = GraphedStep(r, loss, G, view=0, views_in_flight=V, profile=True)
agg = defaultdict(list)
spans = []
for it in range(12):
    for j in range(V):
        v = (it * V + j) % NV
        step.set_inputs(staged[v][0], staged[v][1], None, slot=j)
    step.replay()
    tl = _lib.profile_timeline()
    if it < 2:
        continue
    end = max(s + d for _, _, s, d in tl)
    spans.append(end)
    for name, sid, s, d in tl:
        agg[name].append(d)
    last = tl
print(f"V={V}: step span {np.mean(spans):.1f} us ({np.mean(spans) / V:.1f} us per frame); per-kernel duration under contention (mean us, launches per step):")
tot = 0.0
for name, ds in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    per_step = sum(ds) / len(spans)
    tot += per_step
    print(f"  {name:22s} {np.mean(ds):8.1f} us x {len(ds) // len(spans):3d} = {per_step:8.1f} us per step")
print(f"  sum of kernel durations {tot:.1f} us per step = {tot / np.mean(spans):.2f} kernels in flight on average")
# concurrency profile of the last replay
ev = sorted([(s, 1) for _, _, s, d in last] + [(s + d, -1) for _, _, s, d in last])
level, t_prev, hist = 0, 0.0, defaultdict(float)
for t, dlt in ev:
    hist[level] += t - t_prev
    level += dlt
    t_prev = t
span = max(s + d for _, _, s, d in last)
print("  time with k library kernels in flight: " + ", ".join(f"{k}: {100 * v / span:.0f}%" for k, v in sorted(hist.items())))
print("  last replay, per stream (start us: kernel duration us):")
for sid in sorted({x[1] for x in last}):
    row = sorted([(s, n, d) for n, i, s, d in last if i == sid])
    print(f"   stream {sid}: " + " | ".join(f"{s:.0f}:{n.replace('radix_', 'r_').replace('blend_', 'b_')[:10]} {d:.0f}" for s, n, d in row))
