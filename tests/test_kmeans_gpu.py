"""GPU parity: cuvsKMeans{Fit,Predict,ClusterCost} vs the oracle's assignment step (argmin of expanded L2,
ties -> smaller index, unfused_distance_nn.cuh:41-83) — mirrors python/cuvs/cuvs/tests/test_kmeans.py."""
import numpy as np
import pytest
import torch

import oracle
from tests.util import clustered, uniform

pytestmark = pytest.mark.gpu


def _km():
    from cuvs_b200.cluster import kmeans
    return kmeans


@pytest.mark.parametrize("n,d,k", [(10000, 64, 17), (5000, 128, 300), (3000, 5, 2), (70000, 96, 1024), (8000, 300, 50)])
def test_predict_matches_oracle_assignment(n, d, k):
    km = _km()
    x = uniform(n, d, 3)
    c = uniform(k, d, 4)
    labels, inertia = km.predict(km.KMeansParams(n_clusters=k), torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda())
    ol, od = oracle.kmeans_assign(x, c)
    labels = labels.cpu().numpy()
    diff = labels != ol
    # fp32-grade split products: any disagreement must be a numerical tie between two centroids
    if diff.any():
        d_ours = ((x[diff] - c[labels[diff]]) ** 2).sum(1)
        np.testing.assert_allclose(d_ours, od[diff], rtol=1e-4, atol=1e-4)
    assert diff.mean() < 1e-3
    assert inertia == pytest.approx(float(od.astype(np.float64).sum()), rel=1e-4)
    assert km.cluster_cost(torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda()) == pytest.approx(inertia, rel=1e-6)


def test_fit_recovers_separated_clusters():
    km = _km()
    x, centers = clustered(20000, 32, 7, n_centers=8, sigma=0.05)
    cent, inertia, n_iter = km.fit(km.KMeansParams(n_clusters=8, max_iter=50), torch.from_numpy(x).cuda())
    cent = cent.cpu().numpy()
    # every true centre has a fitted centroid nearby, and the cost is the within-cluster variance
    dmat = ((centers[:, None, :] - cent[None, :, :]) ** 2).sum(-1)
    assert (dmat.min(axis=1) < 0.05).all()
    assert inertia == pytest.approx(20000 * 32 * 0.05 ** 2, rel=0.1)
    assert 1 <= n_iter <= 50
    labels, cost = km.predict(km.KMeansParams(n_clusters=8), torch.from_numpy(x).cuda(), torch.from_numpy(cent).cuda())
    assert cost == pytest.approx(inertia, rel=1e-3)


def test_fit_from_given_centroids_is_one_lloyd_step():
    km = _km()
    x = uniform(4000, 16, 9)
    c0 = x[:5].copy()
    cent, _, _ = km.fit(km.KMeansParams(n_clusters=5, init_method="array", max_iter=1), torch.from_numpy(x).cuda(),
                        centroids=torch.from_numpy(c0.copy()).cuda())
    ol, _ = oracle.kmeans_assign(x, c0)
    expect = np.stack([x[ol == j].mean(0) for j in range(5)])
    np.testing.assert_allclose(cent.cpu().numpy(), expect, rtol=1e-4, atol=1e-4)
