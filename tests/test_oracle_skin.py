"""oracle/skin_ref.py against golden vectors produced by the reference's own skinning_weights_from_voxel_grid."""
import os

import numpy as np
import pytest

from helpers import GOLDEN
from oracle import skin_ref

G = np.load(os.path.join(GOLDEN, "skin_golden.npz"))


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_skin_oracle_matches_reference(name):
    a = {k[len(name) + 1:]: G[k] for k in G.files if k.startswith(name + "_")}
    w, _ = skin_ref.skin_weights(a["xyz"], a["center"], a["scale"], a["grid"])
    np.testing.assert_allclose(w, a["w"], rtol=0, atol=2e-6)
    assert np.allclose(w.sum(-1), 1.0, atol=1e-6)
    g_xyz, g_grid = skin_ref.skin_weights_backward(a["xyz"], a["center"], a["scale"], a["grid"], a["gout"])
    assert np.abs(g_xyz - a["g_xyz"]).max() <= 2e-5 * np.abs(a["g_xyz"]).max()
    assert np.abs(g_grid - a["g_grid"]).max() <= 2e-5 * np.abs(a["g_grid"]).max()
