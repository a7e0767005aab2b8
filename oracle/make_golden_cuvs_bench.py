#!/usr/bin/env python
"""Golden vectors from the REFERENCE ITSELF, executed in the build container (test infrastructure).

The reference ships one CPU implementation of the hot path's semantics: the exact k-nearest-neighbour search its benchmark
package uses to generate ground truth when no GPU is present,
    python/cuvs_bench/cuvs_bench/generate_groundtruth/__main__.py:104-171   cpu_search(dataset, queries, k, metric)
    python/cuvs_bench/cuvs_bench/generate_groundtruth/__main__.py:174-214   calc_truth (row batches + k-way merge)
(pure numpy; `metric` is spelled 'squeclidean' there).  This script imports that module from /root/reference, runs it on small
seeded inputs and writes its OUTPUTS to tests/golden/cuvs_bench_cpu_groundtruth.json.  The inputs are not stored: tests
regenerate them from the seeds below with the same numpy Generator calls (`inputs()`), so the fixture stays small.
/root/reference does not exist on the GPU box — the fixture travels, this script does not need to.

    python oracle/make_golden_cuvs_bench.py            # rewrites the fixture
"""
import importlib
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PKG = "/root/reference/python/cuvs_bench"
OUT = os.path.join(ROOT, "tests", "golden", "cuvs_bench_cpu_groundtruth.json")

CASES = [
    # BASELINE configs[0]'s shape: 10k x 128 f32, k = 10
    dict(name="c0_10k_x_128_l2", n=10_000, d=128, nq=100, k=10, metric="squeclidean", seed=123, via="calc_truth"),
    dict(name="small_ip", n=3_000, d=33, nq=50, k=7, metric="inner_product", seed=5, via="cpu_search"),
    dict(name="small_l2_k1", n=777, d=8, nq=40, k=1, metric="squeclidean", seed=9, via="cpu_search"),
]


def inputs(case):
    """Seeded inputs of a case — the SAME calls are made by the tests (tests/test_oracle_golden.py)."""
    rng = np.random.default_rng(case["seed"])
    ds = rng.standard_normal((case["n"], case["d"]), dtype=np.float32)
    qs = rng.standard_normal((case["nq"], case["d"]), dtype=np.float32)
    return ds, qs


def main():
    sys.path.insert(0, REF_PKG)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = importlib.import_module("cuvs_bench.generate_groundtruth.__main__")
    assert not ref.gpu_system and ref.xp.__name__ == "numpy", "expected the reference's numpy fallback path"
    out = {"source": "rapidsai/cuvs python/cuvs_bench/cuvs_bench/generate_groundtruth/__main__.py (cpu_search / calc_truth), "
                     "executed by oracle/make_golden_cuvs_bench.py in the build container",
           "numpy": np.__version__, "cases": []}
    for c in CASES:
        ds, qs = inputs(c)
        if c["via"] == "calc_truth":
            d, i = ref.calc_truth(ds, qs, c["k"], metric=c["metric"])
        else:
            d, i = ref.cpu_search(ds, qs, c["k"], metric=c["metric"])
        out["cases"].append(dict(c, ids=np.asarray(i).astype(int).tolist(),
                                 distances=[[float(np.float32(v)) for v in row] for row in np.asarray(d)]))
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
