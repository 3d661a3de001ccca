"""Import the reference's own Python (read-only, /root/reference) inside this CPU container.

Used only by the golden-vector generators in this directory; never at test/bench run time
(/root/reference does not exist on the GPU box).

The reference modules top-level-import packages that are not installed here (pymeshlab,
taichi, skimage, pytorch_lightning, hydra, easydict, lpips, trimesh, ... and the two CUDA
submodules) and hard-code ``device="cuda"`` in a few tensor constructors
(src/utils/gaussian_utils.py:249,279,305).  We (1) satisfy the missing imports with inert
stub modules and (2) redirect "cuda" tensor constructors to the CPU.  No reference arithmetic is
replaced: every number in the golden files is computed by the reference's code.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import sys
import types

REF_ROOT = "/root/reference"


class _Anything:
    """Inert object: callable, attribute-able, usable as decorator / base class factory."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]  # decorator use
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name[:1].isupper():
            cls = type(name, (), {"__init__": lambda self, *a, **k: None,
                              "__getattr__": lambda self, n: _Anything()})
            setattr(self, name, cls)
            return cls
        return _Anything()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, roots):
        self.roots = set(roots)

    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_MISSING = [
    "pymeshlab", "diff_gaussian_rasterization", "skimage", "taichi", "simple_knn", "omegaconf",
    "pytorch_lightning", "hydra", "easydict", "lpips", "trimesh", "pysdf", "h5py", "natsort",
    "matplotlib", "plotly", "termcolor", "imageio", "open3d", "kornia", "torchvision", "lightning",
]


def install():
    import torch

    really_missing = []
    for name in _MISSING:
        if name in sys.modules:
            continue
        try:
            if importlib.util.find_spec(name) is None:
                really_missing.append(name)
        except (ImportError, ValueError):
            really_missing.append(name)
    sys.meta_path.append(_StubFinder(really_missing))
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    # device="cuda" -> CPU for the handful of constructors the reference hard-codes.
    def _cpu(fn):
        def wrapped(*a, **k):
            dev = k.get("device")
            if dev is not None and "cuda" in str(dev):
                k["device"] = "cpu"
            return fn(*a, **k)
        return wrapped

    for fname in ("zeros", "ones", "empty", "tensor", "eye", "zeros_like", "ones_like"):
        setattr(torch, fname, _cpu(getattr(torch, fname)))
    return really_missing


def easydict_passthrough():
    """hand_dynamic.forward wraps its result in easydict.EasyDict; give the stub a dict subclass."""
    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__(d or {}, **kw)
            self.__dict__ = self

    importlib.import_module("easydict").EasyDict = EasyDict
    return EasyDict
