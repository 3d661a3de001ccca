"""Oracle for ``simple_knn._C.distCUDA2`` (TEST INFRASTRUCTURE -- see oracle/__init__.py).

PARITY UNPINNED: gitlab.inria.fr/bkerbl/simple-knn is cloned un-pinned at install time
(/root/reference/setup_env.sh:7,12-13) and absent from /root/reference; its only call site is
/root/reference/src/models/gaussian.py:110.  Published behaviour: for every point, the mean of
the squared distances to its 3 nearest *other* points (exact; duplicates give 0).
"""
from __future__ import annotations

import numpy as np


def dist2_knn3(points: np.ndarray) -> np.ndarray:
    """Exact answer with a float64 k-d tree on the float32 inputs (O(N log N))."""
    from scipy.spatial import cKDTree

    p = np.asarray(points, dtype=np.float32).astype(np.float64)
    n = p.shape[0]
    k = min(4, n)
    d, _ = cKDTree(p).query(p, k=k)
    d = d.reshape(n, k)[:, 1:]
    out = np.zeros(n, np.float64)
    if d.shape[1]:
        # upstream always divides by 3; missing neighbours (N < 4) contribute FLT_MAX in upstream -- not exercised by MANUS
        out = (d ** 2).sum(1) / 3.0
    return out.astype(np.float32)


def dist2_knn3_bruteforce(points: np.ndarray) -> np.ndarray:
    """O(N^2) float32 restatement for small N: squared distances accumulated in float32 like a GPU kernel would."""
    p = np.asarray(points, dtype=np.float32)
    n = p.shape[0]
    out = np.zeros(n, np.float32)
    for i in range(n):
        diff = p - p[i]
        d2 = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1] + diff[:, 2] * diff[:, 2]).astype(np.float32)
        d2[i] = np.inf
        best = np.sort(d2)[:3]
        out[i] = np.float32((best[0] + best[1] + best[2]) / np.float32(3.0))
    return out


def contact_dist(pt1: np.ndarray, pt2: np.ndarray):
    """Restates get_contact_dist (/root/reference/src/utils/gaussian_utils.py:521-554): for every point of pt1 the
    Euclidean distance to the nearest point of pt2 and its index, first index of the minimum (strict '<' in ascending j).
    The reference loop is a taichi kernel (taichi is not installable here), so this restatement is pinned only by the
    exact k-d tree cross-check in tests/test_oracle_knn_contact.py.  O(N*M) in float32, chunked."""
    a, b = np.asarray(pt1, np.float32), np.asarray(pt2, np.float32)
    dist, idx = np.zeros(a.shape[0], np.float32), np.zeros(a.shape[0], np.int64)
    for s in range(0, a.shape[0], 2048):
        d = a[s:s + 2048, None, :] - b[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]      # accumulation order of the reference loop
        dd = np.sqrt(d2)
        idx[s:s + 2048] = dd.argmin(1)                                                    # first minimum
        dist[s:s + 2048] = dd.min(1)
    return dist, idx
