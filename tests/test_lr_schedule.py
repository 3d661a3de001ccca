"""The xyz learning-rate schedule (manus_b200.optim.get_expon_lr_func) against values produced by the reference's own
get_expon_lr_func (tests/golden/make_golden_lr.py; src/utils/gaussian_utils.py:212-247)."""
import os

import numpy as np

from manus_b200.optim import get_expon_lr_func

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lr_golden.npz")


def test_schedule_matches_the_reference():
    g = np.load(GOLDEN)
    for name in ("training_setup", "delayed", "disabled"):
        a, b, ds, dm, ms = g[name + "_args"]
        f = get_expon_lr_func(lr_init=float(a), lr_final=float(b), lr_delay_steps=int(ds), lr_delay_mult=float(dm), max_steps=int(ms))
        got = np.array([f(int(s)) for s in g["steps"]])
        np.testing.assert_allclose(got, g[name], rtol=1e-13, atol=0)


def test_schedule_end_points():
    f = get_expon_lr_func(1e-3, 1e-5, max_steps=100)
    assert f(-1) == 0.0 and abs(f(0) - 1e-3) < 1e-18 and abs(f(100) - 1e-5) < 1e-18 and abs(f(10**6) - 1e-5) < 1e-18
    assert abs(f(50) - 1e-4) < 1e-16                                    # log-linear: the geometric mean half way
    assert get_expon_lr_func(0.0, 0.0)(5) == 0.0
