"""The distCUDA2 oracle (oracle_knn_mean_dist2 in oracle/grpg_oracle.c: brute force, the reference's rounding of the
squared distance) against an independent exact nearest-neighbour search (scipy cKDTree, float64)."""
import numpy as np
from scipy.spatial import cKDTree

import knn_cases
from oracle import oracle


def test_oracle_vs_kdtree():
    for name, pts in knn_cases.clouds().items():
        p = pts.numpy().astype(np.float32)
        got = oracle.knn_mean_dist2(p)
        d, _ = cKDTree(p.astype(np.float64)).query(p.astype(np.float64), k=4)  # the first hit is the point itself
        want = (d[:, 1:] ** 2).mean(1)
        assert np.allclose(got, want, rtol=2e-5, atol=1e-12), name
        assert got.dtype == np.float32 and (got >= 0).all()


def test_fewer_than_four_points_keep_the_reference_behaviour():
    """simple_knn.cu:146,183: unfilled slots stay FLT_MAX: the sum is +inf for P < 3 and FLT_MAX / 3 for P = 3"""
    for name, pts in knn_cases.tiny().items():
        got = oracle.knn_mean_dist2(pts.numpy())
        assert got.shape == (pts.shape[0],)
        assert (got > 1e38).all() if pts.shape[0] < 4 else (got < 10).all(), name
