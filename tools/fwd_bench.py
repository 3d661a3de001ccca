#!/usr/bin/env python
"""Kernel time of the forward / backward tile kernels on the headline scene under different environment settings:
python tools/fwd_bench.py "NAME=VALUE,NAME=VALUE" ..."""
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
_lib.profile_enable(True); _lib.profile_report()
for it in range(24):
    out = r.render(it %% 8, sink=r.flat.grads)
    (out["render"].sum()).backward()
    if it == 7: _lib.profile_report()
rep = _lib.profile_report()
print({k: round(ms / n * 1e3, 1) for k, (n, ms) in rep.items() if k.startswith("blend")})
''' % ROOT
for spec in sys.argv[1:] or [""]:
    env = dict(os.environ)
    for kv in filter(None, spec.split(",")):
        if "=" not in kv:
            env["MANUS_B200_LIB"] = os.path.abspath(kv)      # a variant library
            continue
        k, v = kv.split("=", 1)
        env[k] = v
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(spec or "default", out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:], flush=True)
