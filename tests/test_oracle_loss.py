"""The loss oracle (oracle/loss_ref.py) against golden vectors computed by the reference's own loss_utils
(tests/golden/make_golden_loss.py): pins the row-wise SSIM quirk and the 0.8 / 0.2 weighting."""
import os

import numpy as np
import pytest

from helpers import GOLDEN
from oracle import loss_ref

G = np.load(os.path.join(GOLDEN, "loss_golden.npz"))


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_loss_oracle_matches_reference(name, dtype):
    out = loss_ref.photometric_loss(G[f"{name}_pred"], G[f"{name}_gt"], 0.8, 0.2, dtype)
    assert abs(out["l1"] - float(G[f"{name}_l1"])) <= 1e-6
    assert abs(out["ssim"] - float(G[f"{name}_ssim"])) <= 2e-6
    assert abs(out["loss"] - float(G[f"{name}_loss"])) <= 1e-6
    ref = G[f"{name}_grad"]
    assert np.abs(out["grad"] - ref).max() <= 2e-6 * np.abs(ref).max()      # measured: 4e-7 (fp32 summation order)


def test_window_is_separable_like_the_reference():
    g, M = loss_ref.window_factors()
    assert abs(float(g.sum()) - 1.0) < 1e-6 and np.allclose(M, M.T)
    assert M[0, 0] == g[5] and M[0, 2] == g[7] and M[2, 0] == g[3]
