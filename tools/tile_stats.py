#!/usr/bin/env python
"""Per-tile workload statistics of the headline scene (list length, consumed length = deepest last contributor)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manus_b200 import rasterizer as rz, synth  # noqa: E402
from manus_b200.dist import SceneRenderer  # noqa: E402
from manus_b200.pose import pose_gaussians  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
dev = torch.device("cuda", 0)
scene = synth.make_composite(n, seed=0)
r = SceneRenderer(scene, dev, 1920, 1080)
for view in (0, 10, 25, 40):
    cam, c, b = r.view_inputs_host(view)
    c, b = c.to(dev), b.to(dev)
    with torch.no_grad():
        bone_tf = torch.cat([torch.bmm(b.view(-1, 4, 4), r.rest_inv), r._eye], 0)
        px, pc, col, op = pose_gaussians(*[p.detach() for p in r.flat.leaves()], r.skin, bone_tf, c[32:35], 3, False, r.n_hand)
        s = rz.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, r.bg, 1.0, c[0:16], c[16:32], 3, c[32:35], False, False)
        _, radii, st = rz.rasterize_forward(s, px, op.reshape(-1), colors_precomp=col, cov3D_precomp=pc)
        dv = rz.debug_views(st)
        ranges = dv["ranges"].cpu().numpy().astype(np.int64)
        ln = ranges[:, 1] - ranges[:, 0]
        ml = dv["tile_maxlast"].cpu().numpy().astype(np.int64)
        nc = dv["n_contrib"].cpu().numpy()
        q = lambda a: [int(np.percentile(a[a > 0], p)) for p in (50, 90, 99, 99.9, 100)] if (a > 0).any() else []
        print(f"view {view}: D={ln.sum()} active tiles={int((ln > 0).sum())}/{ln.size} list len p50/90/99/99.9/max={q(ln)} | "
              f"consumed (maxlast) sum={ml.sum()} p50/90/99/99.9/max={q(ml)} | tiles with maxlast>512: {(ml > 512).sum()}, >1024: {(ml > 1024).sum()}, "
              f">2048: {(ml > 2048).sum()} | mean n_contrib over covered px={nc[nc > 0].mean():.1f} radii p50/99={np.percentile(radii.cpu().numpy(), [50, 99])}")
