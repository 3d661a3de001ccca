#!/usr/bin/env python
"""Per-kernel ms of one forward+backward frame on the headline scene for library variants / environment settings:
python tools/frame_bench.py "" path/to/variant.so "NAME=VALUE" ..."""
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
G = torch.rand(1080, 1920, 3, device="cuda")
_lib.profile_enable(True); _lib.profile_report()
for it in range(32):
    out = r.render(it %% 8, sink=r.flat.grads)
    (out["render"] * G).sum().backward()
    if it == 7: _lib.profile_report()
rep = _lib.profile_report()
tot = sum(ms for _, ms in rep.values()) / 24
print(round(tot, 4), {k: round(ms / 24, 4) for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])})
''' % ROOT
for spec in sys.argv[1:] or [""]:
    env = dict(os.environ)
    for kv in filter(None, spec.split(",")):
        if "=" not in kv:
            env["MANUS_B200_LIB"] = os.path.abspath(kv)
            continue
        k, v = kv.split("=", 1)
        env[k] = v
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(spec or "default", out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:], flush=True)
