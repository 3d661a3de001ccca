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
