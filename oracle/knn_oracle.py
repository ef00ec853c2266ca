"""CPU oracle for the initial-scale statistic: mean squared distance from every point to its three nearest OTHER points
(what scene/gaussian_model.py:159 obtains from simple_knn._C.distCUDA2).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: the simple-knn submodule (gitlab.inria.fr/bkerbl/simple-knn, .gitmodules:1-3) is
absent from /root/reference and the reference holds no test or golden vector for it, so the definition below restates the
published behaviour [UNVERIFIED_VS_REFERENCE]: squared Euclidean distances, self excluded by index (coincident points count
as distance 0), mean of the three smallest.  With fewer than three other points: mean over those that exist, 0 for one point.
float64 brute force for small inputs, scipy's cKDTree beyond."""
from __future__ import annotations

import numpy as np


def mean_dist2_bruteforce(points: np.ndarray) -> np.ndarray:
    p = np.asarray(points, dtype=np.float32).astype(np.float64)        # the inputs are float32; the arithmetic is exact-ish
    n = p.shape[0]
    if n == 0:
        return np.zeros(0)
    d = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
    d[np.arange(n), np.arange(n)] = np.inf
    k = min(3, n - 1)
    if k == 0:
        return np.zeros(n)
    return np.sort(d, axis=1)[:, :k].mean(axis=1)


def mean_dist2_kdtree(points: np.ndarray) -> np.ndarray:
    from scipy.spatial import cKDTree
    p = np.asarray(points, dtype=np.float32).astype(np.float64)
    n = p.shape[0]
    if n < 5:
        return mean_dist2_bruteforce(points)
    dist, idx = cKDTree(p).query(p, k=4)
    out = np.empty(n)
    for i in range(n):                                                 # drop exactly one entry: the point itself
        row, who = dist[i], idx[i]
        mine = np.nonzero(who == i)[0]
        drop = mine[0] if len(mine) else int(np.argmin(row))           # hidden among coincident points: any zero is the same
        out[i] = (np.delete(row, drop)[:3] ** 2).mean()
    return out
