#!/usr/bin/env python
"""Kernel times of the pose kernels on the headline scene for one or more library builds (CUDA events inside the library)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from manus_b200 import _lib, synth
from manus_b200.dist import SceneRenderer
scene = synth.make_composite(500_000, seed=0)
r = SceneRenderer(scene, torch.device("cuda", 0), 1920, 1080)
from manus_b200 import rasterizer as rz
_lib.profile_enable(True); _lib.profile_report()
for it in range(12):
    out = r.render(it %% 4, sink=r.flat.grads)
    (out["render"].sum()).backward()
rep = _lib.profile_report()
print({k: round(ms / n * 1e3, 1) for k, (n, ms) in rep.items() if k.startswith("pose")})
''' % ROOT
for spec in sys.argv[1:] or [""]:
    env = dict(os.environ)
    lib, *extra = spec.split(",")          # "path/to/lib.so,NAME=VALUE,..." sets environment variables for that run
    for kv in extra:
        k, v = kv.split("=", 1)
        env[k] = v
    if lib:
        env["MANUS_B200_LIB"] = os.path.abspath(lib)
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(os.path.basename(spec) or "default", out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:])
