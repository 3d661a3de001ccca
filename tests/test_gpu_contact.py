"""Contact distance (nearest reference point) kernel against the oracle restatement of get_contact_dist and an exact
k-d tree at the composite scene's size."""
import numpy as np
import pytest
import torch
from scipy.spatial import cKDTree

from oracle import knn_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m", [(1, 1), (5, 3), (1000, 1025), (4097, 700)])
def test_contact_matches_oracle(built_lib, n, m):
    from manus_b200.knn import get_contact_dist, get_contact_map

    rng = np.random.default_rng(n * 7 + m)
    a, b = rng.normal(0, 0.05, (n, 3)).astype(np.float32), rng.normal(0.01, 0.05, (m, 3)).astype(np.float32)
    if m > 4:
        b[3] = b[1]                                       # duplicates: lowest index
        a[0] = b[3]
    d_ref, i_ref = knn_ref.contact_dist(a, b)
    d, i = get_contact_dist(torch.tensor(a, device="cuda"), torch.tensor(b, device="cuda"))
    assert i.dtype == torch.float32                       # the reference returns the indices as float32
    np.testing.assert_allclose(d.cpu().numpy(), d_ref, rtol=1e-6, atol=1e-9)
    np.testing.assert_array_equal(i.cpu().numpy().astype(np.int64), i_ref)
    np.testing.assert_allclose(get_contact_map(torch.tensor(a, device="cuda"), torch.tensor(b, device="cuda")).cpu().numpy(), d_ref, rtol=1e-6, atol=1e-9)


def test_contact_composite_size_vs_kdtree(built_lib):
    """300k hand points against 200k object points."""
    from manus_b200 import synth
    from manus_b200.knn import nearest_point

    sc = synth.make_composite(500_000, seed=0)
    a, b = sc.xyz[: sc.n_hand], sc.xyz[sc.n_hand:]
    d, i = nearest_point(torch.tensor(a, device="cuda"), torch.tensor(b, device="cuda"))
    dk, ik = cKDTree(b.astype(np.float64)).query(a.astype(np.float64), workers=-1)
    np.testing.assert_allclose(d.cpu().numpy(), dk, rtol=3e-6, atol=1e-9)
    assert (i.cpu().numpy() == ik).mean() > 0.9999


def test_contact_against_the_reference_get_contact_map(built_lib):
    """Golden from the reference's own get_contact_map (tests/golden/make_golden_contact.py).  The kernel is exact; the
    reference's torch.cdist values carry up to 6e-5 of their own error (see test_oracle_knn_contact.py)."""
    import os

    from manus_b200.knn import get_contact_dist, get_contact_map

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "contact_golden.npz"))
    a, b = torch.tensor(g["pt1"], device="cuda"), torch.tensor(g["pt2"], device="cuda")
    cm = get_contact_map(a, b, chunk=1024).cpu().numpy()
    np.testing.assert_allclose(cm, g["dist64"], rtol=3e-6, atol=1e-9)
    np.testing.assert_allclose(cm, g["contact_map"], rtol=0, atol=1e-4)
    d, i = get_contact_dist(a, b)
    np.testing.assert_array_equal(d.cpu().numpy(), cm)
    assert (i.cpu().numpy().astype(np.int64) == g["idx64"]).mean() > 0.999
