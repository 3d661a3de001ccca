#!/usr/bin/env python
"""Kernel time of mb_sh_grad_from_views for R = 1, 2, 4, 8 views on the headline scene (one GPU, synthetic per-view data)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manus_b200 import _lib, synth  # noqa: E402
from manus_b200.dist import SceneRenderer  # noqa: E402
from manus_b200.pose import sh_grad_from_views  # noqa: E402

scene = synth.make_composite(500_000, seed=0)
r = SceneRenderer(scene, torch.device("cuda", 0), 1920, 1080)
n = scene.n
for R in (1, 2, 4, 8):
    rec = n * 3 + 21 * 16 + 4
    buf = torch.randn(R, rec, device="cuda") * 0.01
    for v in range(R):
        _, c, b = r.view_inputs_host(v)
        bt = torch.eye(4, device="cuda").repeat(21, 1, 1)
        torch.bmm(b.cuda().view(-1, 4, 4), r.rest_inv, out=bt[:20])
        buf[v, n * 3: n * 3 + 336] = bt.reshape(-1)
        buf[v, n * 3 + 336: n * 3 + 339] = c.cuda()[32:35]
    g = buf[:, : n * 3].unflatten(1, (n, 3))
    bones = buf[:, n * 3: n * 3 + 336].unflatten(1, (21, 4, 4))
    cam = buf[:, n * 3 + 336: n * 3 + 339]
    _lib.profile_enable(True)
    _lib.profile_report()
    for _ in range(10):
        sh_grad_from_views(r.flat.params["xyz"], r.skin, r.n_hand, 3, 16, bones, cam, g, r.flat.grads["f_dc"], r.flat.grads["f_rest"])
    print(R, {k: round(ms / k2 * 1e3, 1) for k, (k2, ms) in _lib.profile_report().items()})
