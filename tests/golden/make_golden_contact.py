"""Golden vectors for the contact map FROM THE REFERENCE'S OWN PYTHON (needs /root/reference):
    python tests/golden/make_golden_contact.py          -> tests/golden/contact_golden.npz
Runs src/utils/gaussian_utils.py:514-518 ``get_contact_map`` (chunked ``torch.cdist(pt1, pt2).min(1)``), the pure-torch twin of
the taichi loop ``get_contact_dist`` (:521-554, cannot run here: taichi is not installed), on hand-like and object-like point
sets (a hand shell 3-12 mm above / inside an object surface, like the grasps of src/modules/composite.py:151-175).
``torch.cdist`` evaluates |a|^2 + |b|^2 - 2ab for sets this large, so its values carry ~1e-7 absolute error; the file also
stores the float64 brute-force distances for reference."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import as R  # noqa: E402

R.install()
import torch  # noqa: E402

import src.utils.gaussian_utils as ref_gu  # noqa: E402

torch.set_num_threads(4)


def main():
    rng = np.random.default_rng(12)
    m, n = 2500, 3100
    d = rng.standard_normal((m, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    obj = (np.array([0.0, 0.0, 0.1]) + d * np.array([0.055, 0.035, 0.045])).astype(np.float32)          # object surface
    h = rng.standard_normal((n, 3)); h /= np.linalg.norm(h, axis=1, keepdims=True)
    h[:, 2] = np.abs(h[:, 2])                                                                            # a shell over the upper half
    hand = (np.array([0.0, 0.0, 0.1]) + h * np.array([0.055, 0.035, 0.045]) * rng.uniform(0.93, 1.25, (n, 1))).astype(np.float32)
    hand[:40] = obj[rng.integers(0, m, 40)]                                                              # exact contacts (distance 0)
    cm = ref_gu.get_contact_map(torch.tensor(hand), torch.tensor(obj), chunk=1024).numpy()
    d64 = np.sqrt(((hand[:, None, :].astype(np.float64) - obj[None].astype(np.float64)) ** 2).sum(-1))
    out = dict(pt1=hand, pt2=obj, contact_map=cm, dist64=d64.min(1), idx64=d64.argmin(1).astype(np.int64))
    path = os.path.join(HERE, "contact_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "max |ref - exact| =", float(np.abs(cm - out["dist64"]).max()), "min dist", float(out["dist64"][40:].min()))


if __name__ == "__main__":
    main()
