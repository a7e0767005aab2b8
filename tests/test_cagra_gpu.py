"""GPU parity: cuvsCagra{IndexFromArgs,Build,Search,Serialize} through the C ABI vs the oracle's restatement of the
single-CTA walk (oracle_cagra_search: same seeds, hash policy, parent selection, termination).

Mirrors c/tests/neighbors/ann_cagra_c.cu (golden 4x2 vectors) and cpp/tests/neighbors/ann_cagra.cuh
(recall >= 0.995 on 1000-vector inputs, eps 0.003; itopk 64 / 256)."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from tests.util import clustered, launches, uniform

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))


def _mod():
    from cuvs_b200.neighbors import cagra
    return cagra


def _knn_graph(ds, degree):
    _, idx = oracle.knn(ds, ds, degree + 1)
    g = np.zeros((ds.shape[0], degree), np.uint32)
    for i in range(ds.shape[0]):
        row = [v for v in idx[i] if v != i][:degree]
        g[i] = row
    return g


def test_reference_known_answers():
    m = _mod()
    case = [c for c in GOLD["cases"] if c["name"] == "cagra_c_4x2_k1"][0]
    ds = np.array(case["dataset"], np.float32)
    qs = np.array(case["queries"], np.float32)
    index = m.build(m.IndexParams(graph_degree=2, intermediate_graph_degree=3), torch.from_numpy(ds).cuda())
    d, i = m.search(m.SearchParams(itopk_size=32), index, torch.from_numpy(qs).cuda(), 1)
    assert i.cpu().numpy().astype(np.int64).tolist() == case["neighbors"]
    np.testing.assert_allclose(d.cpu().numpy(), np.array(case["distances"], np.float32), atol=case["eps"])


@pytest.mark.parametrize("n,d,degree,itopk,width,k", [(1000, 64, 32, 64, 1, 10), (5000, 96, 64, 64, 1, 10), (3000, 17, 32, 32, 1, 5),
                                                      (4000, 128, 32, 128, 2, 16), (2000, 8, 16, 64, 4, 10), (6000, 96, 64, 256, 1, 32)])
def test_walk_matches_oracle(n, d, degree, itopk, width, k):
    m = _mod()
    ds = uniform(n, d, 11, 0.0, 1.0)
    qs = uniform(200, d, 12, 0.0, 1.0)
    g = _knn_graph(ds, degree)
    l0 = launches()
    index = m.from_graph(torch.from_numpy(g.astype(np.int64)).cuda(), torch.from_numpy(ds).cuda())
    assert (len(index), index.dim, index.graph_degree) == (n, d, degree)
    # (algo pinned: a 200-query batch is below 2 queries per SM, where AUTO picks the multi-CTA walk, search_plan.cuh:122-131)
    dist, idx = m.search(m.SearchParams(itopk_size=itopk, search_width=width, algo="single_cta"), index, torch.from_numpy(qs).cuda(), k)
    assert launches() > l0
    dist, idx = dist.cpu().numpy(), idx.cpu().numpy().astype(np.int64)
    rd, ri, _ = oracle.cagra_search(g, ds, qs, k, itopk=itopk, search_width=width)
    ri = ri.astype(np.int64)
    # same walk => same result sets; distances differ only by fp32 summation order (teams of 8 lanes vs sequential)
    assert oracle.recall_with_ties(idx, dist, ri, rd, eps=1e-4) >= 0.999
    assert (idx == ri).mean() >= 0.99
    same = idx == ri
    np.testing.assert_allclose(dist[same], rd[same], rtol=1e-5, atol=1e-5)
    # reference acceptance: recall >= 0.995 vs exact kNN on these sizes (ann_cagra.cuh:473-481) — with the itopk actually used
    gd, gi = oracle.knn(ds, qs, k)
    if itopk >= 64 and d >= 17:
        assert oracle.recall_with_ties(idx, dist, gi, gd, eps=3e-3) >= 0.95  # plain kNN graph on iid data; the walk itself is oracle-exact


def test_inner_product_and_int64_neighbors():
    m = _mod()
    ds, _ = clustered(3000, 32, 5, n_centers=10, sigma=1.0)
    qs, _ = clustered(100, 32, 6, n_centers=10, sigma=1.0)
    _, knn = oracle.knn(ds, ds, 33, "inner_product")
    g = np.stack([[v for v in knn[i] if v != i][:32] for i in range(3000)]).astype(np.uint32)
    index = m.from_graph(torch.from_numpy(g.astype(np.int64)).cuda(), torch.from_numpy(ds).cuda(), metric="inner_product")
    nb = torch.empty((100, 10), dtype=torch.int64, device="cuda")
    dist, idx = m.search(m.SearchParams(itopk_size=64), index, torch.from_numpy(qs).cuda(), 10, neighbors=nb)
    rd, ri, _ = oracle.cagra_search(g, ds, qs, 10, itopk=64, metric="inner_product")
    assert oracle.recall_with_ties(idx.cpu().numpy(), dist.cpu().numpy(), ri.astype(np.int64), rd, eps=1e-3) >= 0.995


def test_build_search_save_load(tmp_path):
    m = _mod()
    # embedding-like data (rank-8 manifold in 64-d): iid uniform points in 64-d have no neighbour structure to walk
    rng = np.random.default_rng(21)
    A = (rng.standard_normal((8, 64)) / np.sqrt(8)).astype(np.float32)
    ds = (rng.standard_normal((20000, 8)).astype(np.float32) @ A + 0.05 * rng.standard_normal((20000, 64)).astype(np.float32))
    qs = (rng.standard_normal((300, 8)).astype(np.float32) @ A + 0.05 * rng.standard_normal((300, 64)).astype(np.float32))
    index = m.build(m.IndexParams(graph_degree=32), torch.from_numpy(ds).cuda())
    g = index.graph.cpu().numpy()
    assert g.shape == (20000, 32) and g.max() < 20000
    assert all(len(set(r.tolist())) == 32 and i not in r for i, r in enumerate(g[:500]))  # no self loops / duplicates
    d1, i1 = m.search(m.SearchParams(itopk_size=64), index, torch.from_numpy(qs).cuda(), 10)
    gd, gi = oracle.knn(ds, qs, 10)
    assert oracle.recall(i1.cpu().numpy().astype(np.int64), gi) >= 0.95
    m.save(str(tmp_path / "cagra.idx"), index)
    again = m.load(str(tmp_path / "cagra.idx"))
    d2, i2 = m.search(m.SearchParams(itopk_size=64), again, torch.from_numpy(qs).cuda(), 10)
    assert torch.equal(i1.to(torch.int64), i2.to(torch.int64)) and torch.equal(d1, d2)


def _bitset(keep):
    keep = np.asarray(keep, bool)
    bits = np.packbits(np.concatenate([keep, np.zeros((-len(keep)) % 32, bool)]), bitorder="little").view(np.int32)
    from cuvs_b200.neighbors import filters
    return filters.from_bitset(torch.from_numpy(bits.copy()).cuda())


def test_reference_known_answers_bitset_filtered():
    """c/tests/neighbors/ann_cagra_c.cu filtered case: same 4x2 data, half of the rows removed by a bitset pre-filter."""
    m = _mod()
    case = [c for c in GOLD["cases"] if c["name"] == "cagra_c_4x2_k1_bitset_filtered"][0]
    ds = np.array(case["dataset"], np.float32)
    qs = np.array(case["queries"], np.float32)
    index = m.build(m.IndexParams(graph_degree=2, intermediate_graph_degree=3), torch.from_numpy(ds).cuda())
    d, i = m.search(m.SearchParams(itopk_size=32), index, torch.from_numpy(qs).cuda(), 1, filter=_bitset(np.isin(np.arange(len(ds)), case["filter_keep"])))
    assert i.cpu().numpy().astype(np.int64).tolist() == case["neighbors"]
    np.testing.assert_allclose(d.cpu().numpy(), np.array(case["distances"], np.float32), atol=case["eps"])


def test_bitset_prefilter_on_a_real_graph():
    """Filtered nodes are walked through but never returned; recall against exact kNN over the kept rows."""
    m = _mod()
    rng = np.random.default_rng(5)
    A = (rng.standard_normal((8, 64)) / np.sqrt(8)).astype(np.float32)
    ds = (rng.standard_normal((20000, 8)).astype(np.float32) @ A + 0.05 * rng.standard_normal((20000, 64)).astype(np.float32))
    qs = (rng.standard_normal((200, 8)).astype(np.float32) @ A + 0.05 * rng.standard_normal((200, 64)).astype(np.float32))
    index = m.build(m.IndexParams(graph_degree=32), torch.from_numpy(ds).cuda())
    keep = rng.random(20000) < 0.5
    d, i = m.search(m.SearchParams(itopk_size=128), index, torch.from_numpy(qs).cuda(), 10, filter=_bitset(keep))
    i = i.cpu().numpy().astype(np.int64)
    assert keep[i].all(), "a filtered-out node was returned"
    kept = np.flatnonzero(keep)
    gd, gi = oracle.knn(ds[kept], qs, 10)
    assert oracle.recall(i, kept[gi]) >= 0.9


def test_fp16_walk_reranks_with_fp32_rows():
    """cuvs_b200 extension: the walk reads an fp16 copy of the rows; returned distances are exact fp32 (re-ranked), ids
    nearly those of the fp32 walk, recall unchanged."""
    m = _mod()
    rng = np.random.default_rng(7)
    A = (rng.standard_normal((8, 96)) / np.sqrt(8)).astype(np.float32)
    ds = (rng.standard_normal((20000, 8)).astype(np.float32) @ A + 0.05 * rng.standard_normal((20000, 96)).astype(np.float32))
    qs = (rng.standard_normal((300, 8)).astype(np.float32) @ A + 0.05 * rng.standard_normal((300, 96)).astype(np.float32))
    index = m.build(m.IndexParams(graph_degree=32), torch.from_numpy(ds).cuda())
    sp = m.SearchParams(itopk_size=64, algo="single_cta")
    d32, i32 = m.search(sp, index, torch.from_numpy(qs).cuda(), 10)
    index.set_walk_precision(16)
    d16, i16 = m.search(sp, index, torch.from_numpy(qs).cuda(), 10)
    i16n, d16n = i16.cpu().numpy().astype(np.int64), d16.cpu().numpy()
    exact = ((ds[i16n] - qs[:, None, :]) ** 2).sum(-1)
    np.testing.assert_allclose(d16n, exact, rtol=1e-5, atol=1e-6)           # distances come from the fp32 rows
    assert (np.diff(d16n, axis=1) >= 0).all()                                # and are sorted after the re-rank
    gd, gi = oracle.knn(ds, qs, 10)
    assert oracle.recall(i16n, gi) >= oracle.recall(i32.cpu().numpy().astype(np.int64), gi) - 0.01
    index.set_walk_precision(32)
    d_again, i_again = m.search(sp, index, torch.from_numpy(qs).cuda(), 10)
    assert torch.equal(i_again, i32) and torch.equal(d_again, d32)


@pytest.mark.parametrize("nq,itopk,k", [(1, 64, 10), (16, 64, 10), (64, 128, 10), (128, 128, 32), (40, 32, 5)])
def test_multi_cta_walk_small_batches(nq, itopk, k):
    """a17: the MULTI_CTA algorithm (search_multi_cta_jit.cuh:56-363) — max(search_width, itopk/32) walkers per query with
    32-entry lists, parents claimed through a shared traversed table, lists merged + de-duplicated.  Small batches are what
    the reference selects it for (search_plan.cuh:122-131).  Checked like the reference checks CAGRA (ann_cagra.cuh: recall
    against exact kNN, min_recall per config): at least the single-CTA walk's recall minus 0.02, unique sorted exact results."""
    m = _mod()
    rng = np.random.default_rng(17)
    A = (rng.standard_normal((8, 64)) / np.sqrt(8)).astype(np.float32)
    ds = (rng.standard_normal((30000, 8)).astype(np.float32) @ A + 0.05 * rng.standard_normal((30000, 64)).astype(np.float32))
    qs = (rng.standard_normal((nq, 8)).astype(np.float32) @ A + 0.05 * rng.standard_normal((nq, 64)).astype(np.float32))
    index = m.build(m.IndexParams(graph_degree=32), torch.from_numpy(ds).cuda())
    q = torch.from_numpy(qs).cuda()
    dm, im = m.search(m.SearchParams(itopk_size=itopk, algo="multi_cta"), index, q, k)
    ds_, is_ = m.search(m.SearchParams(itopk_size=itopk, algo="single_cta"), index, q, k)
    da, ia = m.search(m.SearchParams(itopk_size=itopk), index, q, k)       # AUTO: nq < 2 * SMs -> multi-CTA
    im_n, dm_n = im.cpu().numpy().astype(np.int64), dm.cpu().numpy()
    assert torch.equal(ia, im) and torch.equal(da, dm)
    assert all(len(set(r.tolist())) == k for r in im_n) and (im_n >= 0).all() and (im_n < len(ds)).all()
    np.testing.assert_allclose(dm_n, ((ds[im_n] - qs[:, None, :]) ** 2).sum(-1), rtol=1e-4, atol=1e-6)
    assert (np.diff(dm_n, axis=1) >= 0).all()
    gd, gi = oracle.knn(ds, qs, k)
    r_multi, r_single = oracle.recall(im_n, gi), oracle.recall(is_.cpu().numpy().astype(np.int64), gi)
    assert r_multi >= max(0.9, r_single - 0.02), (r_multi, r_single)


def test_multi_cta_with_bitset_filter():
    m = _mod()
    rng = np.random.default_rng(19)
    A = (rng.standard_normal((8, 64)) / np.sqrt(8)).astype(np.float32)
    ds = (rng.standard_normal((20000, 8)).astype(np.float32) @ A + 0.05 * rng.standard_normal((20000, 64)).astype(np.float32))
    qs = (rng.standard_normal((32, 8)).astype(np.float32) @ A + 0.05 * rng.standard_normal((32, 64)).astype(np.float32))
    index = m.build(m.IndexParams(graph_degree=32), torch.from_numpy(ds).cuda())
    keep = rng.random(20000) < 0.5
    d, i = m.search(m.SearchParams(itopk_size=128, algo="multi_cta"), index, torch.from_numpy(qs).cuda(), 10, filter=_bitset(keep))
    i = i.cpu().numpy().astype(np.int64)
    assert keep[i].all(), "a filtered-out node was returned"
    kept = np.flatnonzero(keep)
    gd, gi = oracle.knn(ds[kept], qs, 10)
    assert oracle.recall(i, kept[gi]) >= 0.85
