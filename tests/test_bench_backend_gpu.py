"""The cuvs-bench plugin (cuvs_b200/bench_backend.py) on the GPU: the reference's YAML-shaped configs drive build + search of the
four cuVS algorithms through the C ABI; recall is computed by the harness against exact ground truth (oracle)."""
import json

import numpy as np
import pytest

import oracle
from cuvs_b200 import bench_backend as bb

pytestmark = pytest.mark.gpu


def _dataset(n=30000, d=64, nq=400, k=10):
    rng = np.random.default_rng(11)
    A = (rng.standard_normal((12, d)) / np.sqrt(12)).astype(np.float32)
    base = (rng.standard_normal((n, 12)).astype(np.float32) @ A + 0.05 * rng.standard_normal((n, d)).astype(np.float32))
    queries = (rng.standard_normal((nq, 12)).astype(np.float32) @ A + 0.05 * rng.standard_normal((nq, d)).astype(np.float32))
    _, gt = oracle.knn(base, queries, k)
    return bb.Dataset(name="manifold-64-euclidean", training_vectors=base, query_vectors=queries, groundtruth_neighbors=gt,
                      distance_metric="euclidean")


@pytest.mark.parametrize("cfg,floor", [
    ({"name": "cuvs_brute_force", "groups": {"base": {"build": {}, "search": {}}}}, 0.9999),
    ({"name": "cuvs_ivf_flat", "groups": {"base": {"build": {"nlist": [64], "ratio": [1], "niter": [10]}, "search": {"nprobe": [8, 32]}}}}, 0.95),
    ({"name": "cuvs_ivf_pq", "groups": {"base": {"build": {"nlist": [64], "pq_dim": [32], "pq_bits": [8], "ratio": [1], "niter": [10]},
                                                   "search": {"nprobe": [32], "internalDistanceDtype": ["float"],
                                                              "smemLutDtype": ["float", "half", "fp8"], "refine_ratio": [1, 2]}}}}, 0.9),
    ({"name": "cuvs_cagra", "groups": {"base": {"build": {"graph_degree": [32], "intermediate_graph_degree": [64]},
                                                  "search": {"itopk": [64], "search_width": [1]}}}}, 0.95),
], ids=lambda x: x["name"] if isinstance(x, dict) else None)
def test_reference_shaped_configs_run_on_the_library(cfg, floor):
    ds = _dataset()
    recs = bb.run_config(cfg, ds, k=10, batch_size=150, mode="throughput")
    json.dumps(recs)
    searches = [r for r in recs if r["name"].endswith("/search")]
    assert recs[0]["name"] == cfg["name"] + "/build" and recs[0]["success"]
    assert len(searches) >= 1 and all(r["items_per_second"] > 0 for r in searches)
    assert max(r["Recall"] for r in searches) >= floor, [r["Recall"] for r in searches]
