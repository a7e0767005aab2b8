"""float16 / int8 / uint8 datasets and queries through the C ABI's dtype switches (c/src/neighbors/ivf_pq.cpp:80-103,
brute_force.cpp:60-110, ivf_flat.cpp, cagra.cpp:245-264).  The library widens them to fp32 on ingestion, so
the answers must equal those of the fp32 copy of the same values: bit-exact for brute force, recall-equal for IVF / CAGRA."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _data(dtype, n, d, seed):
    rng = np.random.default_rng(seed)
    if dtype == np.float16:
        return rng.uniform(-1, 1, (n, d)).astype(np.float16)
    if dtype == np.int8:
        return rng.integers(-100, 100, (n, d), dtype=np.int8)
    return rng.integers(0, 200, (n, d), dtype=np.uint8)


@pytest.mark.parametrize("dtype", [np.float16, np.int8, np.uint8])
def test_brute_force_narrow_dtypes_match_the_oracle_on_the_widened_values(dtype):
    from cuvs_b200.neighbors import brute_force
    ds, qs = _data(dtype, 5000, 48, 1), _data(dtype, 100, 48, 2)
    index = brute_force.build(torch.from_numpy(ds).cuda())
    d, i = brute_force.search(index, torch.from_numpy(qs).cuda(), 10)
    rd, ri = oracle.knn(ds.astype(np.float32), qs.astype(np.float32), 10)
    assert (i.cpu().numpy() == ri).all() and (d.cpu().numpy() == rd).all()
    if dtype != np.float16:  # (the reference compares the DLPack type CODE only: c/src/neighbors/brute_force.cpp:204)
        with pytest.raises(Exception):
            brute_force.search(index, torch.from_numpy(qs.astype(np.float32)).cuda(), 10)  # type mismatch between index and queries


@pytest.mark.parametrize("dtype", [np.float16, np.int8, np.uint8])
def test_ivf_indexes_narrow_dtypes(dtype):
    from cuvs_b200.neighbors import ivf_flat, ivf_pq
    ds, qs = _data(dtype, 20000, 64, 3), _data(dtype, 100, 64, 4)
    gd, gi = oracle.knn(ds.astype(np.float32), qs.astype(np.float32), 10)
    fl = ivf_flat.build(ivf_flat.IndexParams(n_lists=32, kmeans_n_iters=5), torch.from_numpy(ds).cuda())
    d, i = ivf_flat.search(ivf_flat.SearchParams(n_probes=32), fl, torch.from_numpy(qs).cuda(), 10)
    assert oracle.recall_with_ties(i.cpu().numpy(), d.cpu().numpy(), gi, gd, eps=1e-3) >= 0.999   # all lists probed: exact
    pq = ivf_pq.build(ivf_pq.IndexParams(n_lists=32, pq_dim=32, kmeans_n_iters=5), torch.from_numpy(ds).cuda())
    d, i = ivf_pq.search(ivf_pq.SearchParams(n_probes=32), pq, torch.from_numpy(qs).cuda(), 10)
    f32 = ivf_pq.build(ivf_pq.IndexParams(n_lists=32, pq_dim=32, kmeans_n_iters=5), torch.from_numpy(ds.astype(np.float32)).cuda())
    d2, i2 = ivf_pq.search(ivf_pq.SearchParams(n_probes=32), f32, torch.from_numpy(qs.astype(np.float32)).cuda(), 10)
    # same values in -> the same quality out (the k-means build uses atomics, so two builds are not bit-identical)
    r_narrow, r_f32 = oracle.recall(i.cpu().numpy(), gi), oracle.recall(i2.cpu().numpy(), gi)
    assert abs(r_narrow - r_f32) <= 0.03 and r_narrow >= 0.5, (r_narrow, r_f32)


@pytest.mark.parametrize("dtype", [np.float16, np.uint8])
def test_cagra_narrow_dtypes(dtype):
    from cuvs_b200.neighbors import cagra
    ds, qs = _data(dtype, 8000, 32, 5), _data(dtype, 60, 32, 6)
    index = cagra.build(cagra.IndexParams(graph_degree=32), torch.from_numpy(ds).cuda())
    d, i = cagra.search(cagra.SearchParams(itopk_size=128), index, torch.from_numpy(qs).cuda(), 10)
    gd, gi = oracle.knn(ds.astype(np.float32), qs.astype(np.float32), 10)
    assert oracle.recall_with_ties(i.cpu().numpy().astype(np.int64), d.cpu().numpy(), gi, gd, eps=1e-3) >= 0.9
