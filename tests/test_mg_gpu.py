"""GPU tests of the exports round 1 left untested: cuvsMultiGpuIvfFlat{Build,Search} (single process, all visible GPUs, list-
sharded + grouped ncclAllGather when there is more than one device), cuvsPairwiseDistance, and the library's own exchange
step cuvsB200AllGatherMergeTopK (world size 1 here; world size 2 in tests/test_distributed_nccl.py under torchrun)."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle
from tests.util import clustered, uniform

pytestmark = pytest.mark.gpu


class _MgIndexParams(C.Structure):
    _fields_ = [("base_params", C.c_void_p), ("mode", C.c_int)]


class _MgSearchParams(C.Structure):
    _fields_ = [("base_params", C.c_void_p), ("search_mode", C.c_int), ("merge_mode", C.c_int), ("n_rows_per_batch", C.c_int64)]


@pytest.mark.parametrize("mode", [0, 1])  # CUVS_NEIGHBORS_MG_REPLICATED = 0, CUVS_NEIGHBORS_MG_SHARDED = 1
def test_multi_gpu_ivf_flat_matches_exact_knn(mode):
    """n_probes = n_lists makes IVF-Flat exact within the probed lists = the whole dataset: ids must equal the oracle's exact
    kNN whatever the distribution mode and however many devices the handle spans (1 on the default box, N under --gpus N)."""
    from cuvs_b200._capi import DL, check, lib
    from cuvs_b200.neighbors.ivf_flat import _IndexParamsC, _SearchParamsC
    ds, centers = clustered(30000, 64, 21, n_centers=32)
    qs, _ = clustered(300, 64, 22, centers=centers)
    res = C.c_void_p()
    check(lib.cuvsMultiGpuResourcesCreate(C.byref(res)))
    ip = C.POINTER(_MgIndexParams)()
    sp = C.POINTER(_MgSearchParams)()
    check(lib.cuvsMultiGpuIvfFlatIndexParamsCreate(C.byref(ip)))
    check(lib.cuvsMultiGpuIvfFlatSearchParamsCreate(C.byref(sp)))
    ip.contents.mode = mode
    base = C.cast(ip.contents.base_params, C.POINTER(_IndexParamsC)).contents
    base.n_lists, base.kmeans_n_iters = 64, 10
    C.cast(sp.contents.base_params, C.POINTER(_SearchParamsC)).contents.n_probes = 64
    index = C.c_void_p()
    check(lib.cuvsMultiGpuIvfFlatIndexCreate(C.byref(index)))
    t_ds, t_q = torch.from_numpy(ds), torch.from_numpy(qs)
    nb = torch.empty((300, 10), dtype=torch.int64)
    dd = torch.empty((300, 10), dtype=torch.float32)
    d_ds, d_q, d_nb, d_dd = DL(t_ds), DL(t_q), DL(nb), DL(dd)
    check(lib.cuvsMultiGpuIvfFlatBuild(res, ip, d_ds.ptr, index))
    check(lib.cuvsMultiGpuIvfFlatSearch(res, sp, index, d_q.ptr, d_nb.ptr, d_dd.ptr))
    gd, gi = oracle.knn(ds, qs, 10)
    assert oracle.recall_with_ties(nb.numpy(), dd.numpy(), gi, gd, eps=1e-4) >= 0.999
    assert (nb.numpy() == gi).mean() >= 0.995
    check(lib.cuvsMultiGpuIvfFlatIndexDestroy(index))
    check(lib.cuvsMultiGpuIvfFlatSearchParamsDestroy(sp))
    check(lib.cuvsMultiGpuIvfFlatIndexParamsDestroy(ip))
    check(lib.cuvsMultiGpuResourcesDestroy(res))


@pytest.mark.parametrize("metric,code", [("sqeuclidean", 0), ("euclidean", 1), ("inner_product", 6), ("cosine", 2)])
def test_pairwise_distance(metric, code):
    """cuvsPairwiseDistance vs float64 NumPy (c/tests/distance/pairwise_distance_c.cu shape class: small dense blocks)."""
    from cuvs_b200._capi import DL, check, lib
    from cuvs_b200.common import Resources
    x = uniform(257, 70, 1, -1, 1)
    y = uniform(129, 70, 2, -1, 1)
    out = torch.empty((257, 129), dtype=torch.float32, device="cuda")
    res = Resources()
    xg, yg = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    dx, dy, do = DL(xg), DL(yg), DL(out)  # (held: a DL owns the shape array the DLTensor points at)
    check(lib.cuvsPairwiseDistance(res.get_c_obj(), dx.ptr, dy.ptr, do.ptr, C.c_int(code), C.c_float(2.0)))
    res.sync()
    x64, y64 = x.astype(np.float64), y.astype(np.float64)
    dot = x64 @ y64.T
    sq = (x64 ** 2).sum(1)[:, None] + (y64 ** 2).sum(1)[None, :] - 2 * dot
    ref = {"sqeuclidean": sq, "euclidean": np.sqrt(np.maximum(sq, 0)), "inner_product": dot,
           "cosine": 1 - dot / np.sqrt((x64 ** 2).sum(1)[:, None] * (y64 ** 2).sum(1)[None, :])}[metric]
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=2e-5, atol=2e-5)


def test_allgather_merge_world_of_one():
    """The exchange step inside the library (pack -> ncclAllGather -> merge) with a one-rank communicator: identity."""
    from cuvs_b200._capi import DL, check, lib
    from cuvs_b200.common import Resources
    res = Resources()
    ident = (C.c_ubyte * 128)()
    check(lib.cuvsB200NcclUniqueId(ident))
    comm = C.c_void_p()
    check(lib.cuvsB200CommCreate(res.get_c_obj(), ident, C.c_int(0), C.c_int(1), C.byref(comm)))
    d = torch.sort(torch.rand((500, 10), device="cuda"), dim=1).values
    i = torch.randint(0, 1 << 40, (500, 10), device="cuda")
    od, oi = torch.empty_like(d), torch.empty_like(i)
    h = [DL(d), DL(i), DL(od), DL(oi)]
    check(lib.cuvsB200AllGatherMergeTopK(res.get_c_obj(), comm, h[0].ptr, h[1].ptr, h[2].ptr, h[3].ptr, C.c_bool(True)))
    res.sync()
    assert torch.equal(od, d) and torch.equal(oi, i)
    check(lib.cuvsB200CommDestroy(comm))
