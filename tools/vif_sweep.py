#!/usr/bin/env python
"""Frames/s of the graph-replayed step against the number of views in flight (GraphedStep(views_in_flight=V)), headline
scene, inputs resident.  python tools/vif_sweep.py [V ...]   (a trailing 'o' = ordered accumulation, e.g. 4o)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manus_b200 import rasterizer as rz, synth  # noqa: E402
from manus_b200.dist import GraphedStep, PipelinedStep, SceneRenderer  # noqa: E402

specs = sys.argv[1:] or ["1", "2", "4", "4o", "6", "8"]
W, H, NV = 1920, 1080, 50
dev = torch.device("cuda", 0)
scene = synth.make_composite(500_000, seed=0)
r = SceneRenderer(scene, dev, W, H)
G = torch.rand(3, H, W, device=dev).permute(1, 2, 0)
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
loss_fn = lambda image, target: (torch.dot(image.permute(2, 0, 1).reshape(-1), target.permute(2, 0, 1).reshape(-1)), target)
for spec in specs:
    # suffixes: o = ordered accumulation; p / pN = PipelinedStep with 1 / N ranges (all pose backwards in one multi-view pass)
    if "p" in spec:      # e.g. 4p, 4p2
        V, chunks = int(spec.split("p")[0]), int(spec.split("p")[1] or 1)
        step = PipelinedStep(r, loss_fn, G, view=0, views_in_flight=V, chunks=chunks)
    else:
        V, ordered = int(spec.rstrip("o")), spec.endswith("o")
        step = GraphedStep(r, loss_fn, G, view=0, views_in_flight=V, ordered=ordered)
    K = max(20, 400 // V)

    def run(n, base):
        for it in range(n):
            for j in range(V):
                v = ((base + it) * V + j) % NV
                step.set_inputs(staged[v][0], staged[v][1], None, slot=j)
            step.replay()

    run(5, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(K, 5)
    e1.record()
    torch.cuda.synchronize()
    step.check()
    ms = e0.elapsed_time(e1) / K
    print(f"V={spec}: {ms:.4f} ms per step, {ms / V:.4f} ms per frame, {1e3 * V / ms:.0f} frames/s", flush=True)
    del step
