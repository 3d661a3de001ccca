"""Build the C-ABI shared library (sm_100a only) in-tree:  python -m manus_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  Output: manus_b200/lib/libmanus_b200.so (git-ignored; travels to the GPU box
with the gpurun snapshot).  raster_geom.cu is compiled with --fmad=false (see its header).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libmanus_b200.so")
INCLUDE = os.path.join(HERE, "..", "include")

SOURCES = ["api.cu", "sort_scan.cu", "raster_geom.cu", "raster_blend.cu", "pose.cu", "knn.cu", "loss.cu", "skin.cu", "adam.cu", "exchange.cu", "shcolor.cu"]
PER_FILE_FLAGS = {"raster_geom.cu": ["--fmad=false"]}
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the manus_b200 CUDA library cannot be built (there is no CPU fallback)")


def _deps_mtime() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "manus_b200.h"), __file__]
    return max(os.path.getmtime(f) for f in files)


def build_variant(name: str, flags) -> str:
    """Extra build for kernel experiments: manus_b200/lib/variants/lib<name>.so compiled with additional nvcc flags."""
    out_dir = os.path.join(HERE, "lib", "variants")
    os.makedirs(os.path.join(out_dir, name), exist_ok=True)
    cc = nvcc()
    objs = []
    for src in SOURCES:
        obj = os.path.join(out_dir, name, src.replace(".cu", ".o"))
        cmd = [cc, "-c", os.path.join(CSRC, src), "-o", obj] + ARCH + COMMON + PER_FILE_FLAGS.get(src, []) + list(flags)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stdout + r.stderr)
        objs.append(obj)
    lib = os.path.join(out_dir, f"lib{name}.so")
    r = subprocess.run([cc, "-shared", "-o", lib] + objs + ARCH + ["-Xcompiler", "-fPIC"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stdout + r.stderr)
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    cc = nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [cc, "-c", os.path.join(CSRC, src), "-o", obj] + ARCH + COMMON + PER_FILE_FLAGS.get(src, [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{log}")
        if verbose:
            print(log)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [cc, "-shared", "-o", LIB] + objs + ARCH + ["-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
