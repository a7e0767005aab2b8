"""GPU parity: cuvsBruteForce{Build,Search} (through the C ABI) vs the oracle.

Bit-exact on indices AND distances: the library's last stage re-scores candidates in the oracle's
pinned fp32 arithmetic, and a certificate guarantees the candidate set contains the true top-k
(else the query is recomputed on the exact path).  Mirrors cpp/tests/neighbors/brute_force.cu and
python/cuvs/cuvs/tests/test_brute_force.py.
"""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from tests.util import clustered, last_flagged, launches, uniform

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))


def _bf():
    from cuvs_b200.neighbors import brute_force
    return brute_force


def _search(ds, qs, k, metric="sqeuclidean", prefilter=None):
    bf = _bf()
    index = bf.build(torch.from_numpy(ds).cuda(), metric=metric)
    d, i = bf.search(index, torch.from_numpy(qs).cuda(), k, prefilter=prefilter)
    return d.cpu().numpy(), i.cpu().numpy()


@pytest.mark.parametrize("case", [c for c in GOLD["cases"] if "filter_keep" not in c], ids=lambda c: c["name"])
def test_reference_known_answers(case):
    ds = np.array(case["dataset"], np.float32)
    qs = np.array(case["queries"], np.float32)
    dist, idx = _search(ds, qs, case["k"], case["metric"])
    assert idx.tolist() == case["neighbors"]
    np.testing.assert_allclose(dist, np.array(case["distances"], np.float32), atol=case["eps"])


@pytest.mark.parametrize("case", [c for c in GOLD["cases"] if "filter_keep" in c], ids=lambda c: c["name"])
def test_reference_known_answers_bitset(case):
    from cuvs_b200.neighbors import filters
    ds = np.array(case["dataset"], np.float32)
    qs = np.array(case["queries"], np.float32)
    word = 0
    for j in case["filter_keep"]:
        word |= 1 << j
    f = filters.from_bitset(torch.tensor([word], dtype=torch.int64).to(torch.uint32))
    dist, idx = _search(ds, qs, case["k"], case["metric"], prefilter=f)
    assert idx.tolist() == case["neighbors"]
    np.testing.assert_allclose(dist, np.array(case["distances"], np.float32), atol=case["eps"])


def test_reference_label_case():
    lc = GOLD["label_case"]
    pts = np.array(lc["points"], np.float32)
    labels = np.array(lc["labels"])
    _, idx = _search(pts, pts, lc["k"], lc["metric"])
    assert (labels[idx] == labels[:, None]).all()


@pytest.mark.parametrize("n,d,nq,k", [(20000, 128, 300, 10), (5000, 96, 130, 5), (3001, 33, 77, 1),
                                       (9000, 64, 128, 24), (1000, 128, 5, 10), (70000, 128, 257, 10),
                                       (6000, 200, 140, 10), (4000, 768, 64, 10),  # dims > 128: streamed k-blocks
                                       # 24 < k <= 64 (the reference's fused range): lists of 32 merged to 64 / 96 candidates
                                       (20000, 128, 300, 40), (30000, 128, 200, 64), (3001, 64, 50, 64), (9000, 200, 64, 33)])
@pytest.mark.parametrize("metric", ["sqeuclidean"])
def test_exact_match_uniform(n, d, nq, k, metric):
    ds, qs = uniform(n, d, 1234), uniform(nq, d, 4321)
    l0 = launches()
    dist, idx = _search(ds, qs, k, metric)
    assert launches() > l0, "no kernel of the library was launched"
    rd, ri = oracle.knn(ds, qs, k, metric)
    assert (idx == ri).all(), f"{(idx != ri).sum()} index mismatches"
    assert (dist == rd).all(), f"max |Δ| = {np.abs(dist - rd).max()}"
    # the tensor-core candidate path must carry the result, not the fallback
    assert last_flagged() <= max(1, nq // 100), f"{last_flagged()} of {nq} queries fell back to the exact path"


@pytest.mark.parametrize("metric", ["euclidean", "inner_product", "cosine", "l2_unexpanded", "l2_sqrt_unexpanded"])
def test_exact_match_metrics(metric):
    ds, _ = clustered(12000, 128, 7)
    qs, _ = clustered(200, 128, 8)
    dist, idx = _search(ds, qs, 10, metric)
    rd, ri = oracle.knn(ds, qs, 10, metric)
    assert (idx == ri).all(), f"{(idx != ri).sum()} index mismatches"
    np.testing.assert_array_equal(dist, rd)
    assert last_flagged() <= 4


def test_duplicates_take_the_certified_fallback():
    base = uniform(64, 128, 5)
    ds = np.tile(base, (40, 1))  # every row appears 40 times: exact ties far beyond k'
    qs = uniform(32, 128, 6)
    dist, idx = _search(ds, qs, 10)
    rd, ri = oracle.knn(ds, qs, 10)
    assert last_flagged() == 32
    assert (idx == ri).all() and (dist == rd).all()


@pytest.mark.parametrize("metric", ["inner_product", "cosine", "euclidean"])
def test_k_above_the_fused_list_length(metric):
    """k = 48: the candidates are the merged top-64 of per-(split, column half) lists of 32; the certificate also uses the
    lists' own worst entries, so the result is still the oracle's bit for bit."""
    ds, _ = clustered(15000, 128, 17)
    qs, _ = clustered(150, 128, 18)
    dist, idx = _search(ds, qs, 48, metric)
    rd, ri = oracle.knn(ds, qs, 48, metric)
    assert (idx == ri).all(), f"{(idx != ri).sum()} index mismatches"
    np.testing.assert_array_equal(dist, rd)
    assert last_flagged() <= 8


def test_duplicates_with_large_k_take_the_certified_fallback():
    base = uniform(50, 128, 5)
    ds = np.tile(base, (60, 1))  # 60 copies of every row: more than 32 exact ties inside one candidate list
    qs = uniform(16, 128, 6)
    dist, idx = _search(ds, qs, 40)
    rd, ri = oracle.knn(ds, qs, 40)
    assert (idx == ri).all() and (dist == rd).all()


def _ref_exec_cases():
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cuvs_bench_cpu_groundtruth.json")))["cases"]


@pytest.mark.parametrize("case", _ref_exec_cases(), ids=lambda c: c["name"])
def test_matches_the_reference_cpu_search_outputs(case):
    """Fixtures = outputs of the reference's own CPU exact search (cuvs_bench generate_groundtruth cpu_search / calc_truth, run in
    the build container by oracle/make_golden_cuvs_bench.py) on seeded inputs regenerated here; the first case has BASELINE
    configs[0]'s shape (10k x 128, k = 10)."""
    rng = np.random.default_rng(case["seed"])
    ds = rng.standard_normal((case["n"], case["d"]), dtype=np.float32)
    qs = rng.standard_normal((case["nq"], case["d"]), dtype=np.float32)
    metric = "sqeuclidean" if case["metric"] == "squeclidean" else "inner_product"
    dist, idx = _search(ds, qs, case["k"], metric)
    ri, rd = np.array(case["ids"]), np.array(case["distances"], np.float32)
    np.testing.assert_allclose(dist, rd, rtol=2e-5, atol=1e-5)
    assert (idx == ri).mean() >= 0.999


def test_fewer_rows_than_k():
    ds = np.eye(3, 16, dtype=np.float32)
    dist, idx = _search(ds, ds[:2], 5)
    assert idx[0, :3].tolist() == [0, 1, 2] and idx[0, 3:].tolist() == [-1, -1]


def test_large_k_and_wide_dim_use_exact_path():
    ds, qs = uniform(3000, 200, 1), uniform(20, 200, 2)
    dist, idx = _search(ds, qs, 40)
    rd, ri = oracle.knn(ds, qs, 40)
    assert (idx == ri).all() and (dist == rd).all()


def test_bitmap_filter():
    from cuvs_b200.neighbors import filters
    n, nq = 500, 9
    ds, qs = uniform(n, 32, 3), uniform(nq, 32, 4)
    rng = np.random.default_rng(0)
    keep = rng.random((nq, n)) < 0.3
    bits = np.zeros((nq * n + 31) // 32, np.uint32)
    flat = keep.ravel()
    for pos in np.nonzero(flat)[0]:
        bits[pos >> 5] |= np.uint32(1 << (pos & 31))
    f = filters.from_bitmap(torch.from_numpy(bits.view(np.int32)).view(torch.int32))
    dist, idx = _search(ds, qs, 5, prefilter=f)
    for i in range(nq):
        sub = np.nonzero(keep[i])[0]
        rd, ri = oracle.knn(ds[sub], qs[i:i + 1], 5)
        assert (sub[ri[0]] == idx[i]).all()
        np.testing.assert_array_equal(rd[0], dist[i])


def test_column_major_queries_and_save_load(tmp_path):
    bf = _bf()
    ds, qs = uniform(4000, 64, 11), uniform(50, 64, 12)
    index = bf.build(torch.from_numpy(ds).cuda())
    q_f = torch.from_numpy(qs).cuda().t().contiguous().t()  # F-contiguous view
    assert not q_f.is_contiguous()
    d1, i1 = bf.search(index, q_f, 7)
    rd, ri = oracle.knn(ds, qs, 7)
    assert (i1.cpu().numpy() == ri).all()
    bf.save(str(tmp_path / "bf.idx"), index)
    again = bf.load(str(tmp_path / "bf.idx"))
    d2, i2 = bf.search(again, torch.from_numpy(qs).cuda(), 7)
    assert torch.equal(i1, i2) and torch.equal(d1, d2)


def test_error_paths_do_not_throw_across_the_boundary():
    from cuvs_b200._capi import CuvsError
    bf = _bf()
    index = bf.build(torch.zeros(10, 8, device="cuda"))
    with pytest.raises(CuvsError, match="neighbors should be of type int64_t"):
        bf.search(index, torch.zeros(2, 8, device="cuda"), 3, neighbors=torch.zeros(2, 3, dtype=torch.int32, device="cuda"))
    with pytest.raises(CuvsError, match="device compatible"):
        bf.search(index, torch.zeros(2, 8), 3, neighbors=torch.zeros(2, 3, dtype=torch.int64, device="cuda"),
                  distances=torch.zeros(2, 3, device="cuda"))
    with pytest.raises(CuvsError):
        bf.build(torch.zeros(10, 8, device="cuda"), metric="jaccard")


def test_candidate_error_budget():
    """The certificate's eps (2^-15 (|q|^2 + max|x|^2)) must dominate the observed split-bf16 error."""
    import ctypes as C
    from cuvs_b200._capi import DL, check, lib
    from cuvs_b200.common import Resources
    bf = _bf()
    ds, _ = clustered(30000, 128, 21)
    qs, _ = clustered(256, 128, 22)
    index = bf.build(torch.from_numpy(ds).cuda())
    res = Resources()
    q = torch.from_numpy(qs).cuda()
    pos = torch.zeros(256, 16, dtype=torch.uint32, device="cuda")
    sc = torch.zeros(256, 16, dtype=torch.float32, device="cuda")
    check(lib.cuvsB200BruteForceCandidates(res.get_c_obj(), index._p, DL(q).ptr, DL(pos).ptr, DL(sc).ptr))
    pos = pos.cpu().numpy().astype(np.int64)
    s = sc.cpu().numpy().astype(np.float64)
    qn = (qs.astype(np.float64) ** 2).sum(1)
    approx = qn[:, None] + 2 * s
    exact = ((qs[:, None, :].astype(np.float64) - ds[pos].astype(np.float64)) ** 2).sum(-1)
    xn_max = (ds.astype(np.float64) ** 2).sum(1).max()
    rel = np.abs(approx - exact) / (qn[:, None] + xn_max)
    assert rel.max() < 2.0 ** -15 / 4, f"observed relative error {rel.max():.3e}"
    # and the candidate lists contain the true 10 best (vs float64 ground truth); the tail of the 16 may differ at
    # near-ties because scores carry the 4-bit column tag (<= 2^-19 relative) — that is what the certificate guards
    full = ((qs[:, None, :].astype(np.float64) - ds[None, :, :].astype(np.float64)) ** 2).sum(-1)
    best = np.sort(full, axis=1)[:, :10]
    np.testing.assert_allclose(np.sort(exact, axis=1)[:, :10], best, rtol=1e-9, atol=1e-9)
