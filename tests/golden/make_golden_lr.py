"""Golden values of the xyz learning-rate schedule FROM THE REFERENCE'S OWN PYTHON (needs /root/reference):
    python tests/golden/make_golden_lr.py          -> tests/golden/lr_golden.npz
Runs src/utils/gaussian_utils.py:212-247 ``get_expon_lr_func`` as ``GaussianModel.training_setup`` configures it
(src/models/gaussian.py:143-146; values of config/model/gaussian/gaussian.yaml:4-7 with a spatial_lr_scale of 0.35 folded in) and with a delay."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import as R  # noqa: E402

R.install()
import src.utils.gaussian_utils as ref_gu  # noqa: E402

CASES = {  # name -> (lr_init, lr_final, lr_delay_steps, lr_delay_mult, max_steps)
    "training_setup": (0.0016 * 0.35, 0.0000016 * 0.35, 0, 0.01, 30000),
    "delayed": (2e-3, 1e-5, 500, 0.1, 7000),
    "disabled": (0.0, 0.0, 0, 1.0, 1000),
}
STEPS = np.array([-3, 0, 1, 2, 10, 249, 250, 499, 500, 501, 999, 1000, 6999, 7000, 7001, 15000, 29999, 30000, 30001, 100000])


def main():
    out = {"steps": STEPS}
    for name, (a, b, ds, dm, ms) in CASES.items():
        f = ref_gu.get_expon_lr_func(lr_init=a, lr_final=b, lr_delay_steps=ds, lr_delay_mult=dm, max_steps=ms)
        out[name + "_args"] = np.array([a, b, ds, dm, ms], np.float64)
        out[name] = np.array([f(int(s)) for s in STEPS], np.float64)
    path = os.path.join(HERE, "lr_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v[:4] for k, v in out.items() if not k.endswith("_args") and k != "steps"})


if __name__ == "__main__":
    main()
