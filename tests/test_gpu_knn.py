"""distCUDA2 through the C ABI against the exact k-d tree oracle."""
import numpy as np
import pytest
import torch

from oracle import knn_ref

pytestmark = pytest.mark.gpu


def gpu_knn(p):
    from manus_b200.knn import distCUDA2

    out = distCUDA2(torch.tensor(p, device="cuda"))
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("n", [4, 5, 33, 1000, 1025, 50_000])
def test_random_points(built_lib, n):
    rng = np.random.default_rng(n)
    p = (rng.standard_normal((n, 3)) * np.array([0.05, 0.1, 0.2])).astype(np.float32)
    got, ref = gpu_knn(p), knn_ref.dist2_knn3(p)
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-12)      # fp32 squared distances vs float64 tree


def test_regular_lattice(built_lib):
    h = 0.01
    g = np.stack(np.meshgrid(*[np.arange(12)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * h
    got = gpu_knn(g)
    np.testing.assert_allclose(got, h * h, rtol=1e-4)                 # 3 nearest are axis neighbours at distance h


def test_duplicates_give_zero(built_lib):
    rng = np.random.default_rng(1)
    p = rng.standard_normal((500, 3)).astype(np.float32)
    p = np.concatenate([p, p, p, p], 0)                               # every point has 3 exact copies
    assert np.all(gpu_knn(p) == 0)


def test_hand_scene_matches_oracle_and_bruteforce(built_lib):
    from manus_b200 import synth

    sc = synth.make_hand(30_000, seed=2)
    got = gpu_knn(sc.xyz)
    np.testing.assert_allclose(got, knn_ref.dist2_knn3(sc.xyz), rtol=2e-5, atol=1e-14)
    sub = sc.xyz[:600]
    np.testing.assert_allclose(gpu_knn(sub), knn_ref.dist2_knn3_bruteforce(sub), rtol=1e-5, atol=1e-14)


def test_shim_import(built_lib):
    import os, sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "shims"))
    from simple_knn._C import distCUDA2

    out = distCUDA2(torch.rand(100, 3, device="cuda"))
    assert out.shape == (100,) and out.dtype == torch.float32 and bool((out > 0).all())
