"""The contact-distance oracle (oracle/knn_ref.contact_dist) against an exact float64 k-d tree."""
import numpy as np
from scipy.spatial import cKDTree

from oracle import knn_ref


def test_contact_oracle_matches_kdtree():
    rng = np.random.default_rng(0)
    a, b = rng.normal(0, 0.05, (3000, 3)).astype(np.float32), rng.normal(0.02, 0.04, (5000, 3)).astype(np.float32)
    b[10] = b[4]                                   # duplicate reference point: the first index wins
    a[7] = b[10]
    d, i = knn_ref.contact_dist(a, b)
    dk, ik = cKDTree(b.astype(np.float64)).query(a.astype(np.float64))
    np.testing.assert_allclose(d, dk, rtol=2e-6, atol=1e-9)
    assert i[7] == 4 and d[7] == 0.0
    assert (i == ik).mean() > 0.999


def test_contact_oracle_against_the_reference_get_contact_map():
    """tests/golden/contact_golden.npz: src/utils/gaussian_utils.py:514-518 run here (make_golden_contact.py).  torch.cdist takes
    the |a|^2 + |b|^2 - 2ab route for sets this large, so the reference's own values are off by up to 6e-5 (at exact contacts,
    where the true distance is 0); the oracle is exact, hence the two tolerances."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "contact_golden.npz"))
    d, i = knn_ref.contact_dist(g["pt1"], g["pt2"])
    np.testing.assert_allclose(d, g["dist64"], rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(d, g["contact_map"], rtol=0, atol=1e-4)            # the reference's own cdist noise
    assert float(np.abs(g["contact_map"] - g["dist64"]).max()) < 1e-4
    assert (d[:40] == 0).all()                                                   # the exact contacts
    assert (i == g["idx64"]).mean() > 0.999
