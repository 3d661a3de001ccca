#!/usr/bin/env python
"""BASELINE.json configs 1-4 and the Gaussian-count sweep (config 5) on one GPU: runs bench.py per configuration and
prints / writes a markdown table (frames/s, ms/step, instances, algorithmic GB/s and fraction of the measured HBM peak).
    python tools/sweep.py [out.md]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = [
    ("config 2: static object 100k, 800x800", ["--scene", "object", "--gaussians", "100000", "--width", "800", "--height", "800"]),
    ("config 3: articulated hand 300k, 1080p", ["--scene", "hand", "--gaussians", "300000"]),
    ("config 4 (headline): composite 500k, 1080p", ["--scene", "composite", "--gaussians", "500000"]),
] + [(f"config 5 sweep: composite {n // 1000}k, 1080p", ["--scene", "composite", "--gaussians", str(n)])
     for n in (50_000, 100_000, 200_000, 1_000_000, 2_000_000)]


def main(out=None):
    rows = []
    for name, flags in CONFIGS:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "100", "--warmup", "5", "--no-cpu-baseline"] + flags
        r = subprocess.run(cmd, capture_output=True, text=True)
        line = next((l for l in r.stdout.splitlines() if l.startswith("{")), None)
        if line is None:
            rows.append(f"| {name} | failed: {r.stderr.strip().splitlines()[-1][:120] if r.stderr.strip() else 'no output'} | | | | | | |")
            continue
        d = json.loads(line)
        f = d["frame"]
        top = d["roofline"]
        rows.append(f"| {name} | {d['value']:.0f} | {d['ms_per_step'] / d.get('frames_per_step', 1):.3f} | {d['e2e']['value']:.0f} | {f['num_rendered_mean'] / 1e6:.2f} M | "
                    f"{f['algorithmic_bytes'] / 1e9:.2f} GB | {f['achieved_gbps']:.0f} GB/s ({100 * f['frac_of_hbm_peak']:.0f} %) | "
                    f"{top['kernel']} {top['launch_ms'] * 1e3:.0f} us |")
        print(rows[-1], flush=True)
    txt = ("| configuration | frames/s (resident) | ms/frame | frames/s (e2e, host inputs) | instances D | algorithmic bytes/frame | achieved (of HBM peak) | dominant kernel |\n"
           "|---|---|---|---|---|---|---|---|\n" + "\n".join(rows) + "\n")
    if out:
        with open(out, "w") as fh:
            fh.write("# One-GPU sweep over BASELINE.json's configurations (fwd+bwd, `python tools/sweep.py`)\n\n"
                     "100 timed steps after 5 warm-up steps each, CUDA-graph replay with bench.py's default views in flight, 50 shipped views cycled; algorithmic bytes per SURVEY.md section 8d\n"
                     "with the measured instance count.\n\n" + txt)
    print(txt)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
