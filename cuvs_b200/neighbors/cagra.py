"""CAGRA — same surface as python/cuvs/cuvs/neighbors/cagra/cagra.pyx (IndexParams, Index, build, SearchParams, search,
save, load) plus from_graph() for cuvsCagraIndexFromArgs (the reference exposes it from C/C++ only)."""
from __future__ import annotations

import ctypes as C

import torch

from .._capi import DL, DLManagedTensor, as_tensor, check, cuvsFilter, index_handle, lib, metric_code, view_to_torch
from ..common.resources import auto_sync_resources


class _IndexParamsC(C.Structure):  # include/cuvs/neighbors/cagra.h: struct cuvsCagraIndexParams
    _fields_ = [("metric", C.c_int), ("intermediate_graph_degree", C.c_size_t), ("graph_degree", C.c_size_t),
                ("build_algo", C.c_int), ("nn_descent_niter", C.c_size_t), ("compression", C.c_void_p),
                ("graph_build_params", C.c_void_p)]


class _SearchParamsC(C.Structure):  # struct cuvsCagraSearchParams
    _fields_ = [("max_queries", C.c_size_t), ("itopk_size", C.c_size_t), ("max_iterations", C.c_size_t), ("algo", C.c_int),
                ("team_size", C.c_size_t), ("search_width", C.c_size_t), ("min_iterations", C.c_size_t),
                ("thread_block_size", C.c_size_t), ("hashmap_mode", C.c_int), ("hashmap_min_bitlen", C.c_size_t),
                ("hashmap_max_fill_rate", C.c_float), ("num_random_samplings", C.c_uint32), ("rand_xor_mask", C.c_uint64),
                ("persistent", C.c_bool), ("persistent_lifetime", C.c_float), ("persistent_device_usage", C.c_float)]


class IndexParams:
    def __init__(self, *, metric="sqeuclidean", intermediate_graph_degree=128, graph_degree=64, build_algo="ivf_pq",
                 nn_descent_niter=20):
        self._p = C.POINTER(_IndexParamsC)()
        check(lib.cuvsCagraIndexParamsCreate(C.byref(self._p)))
        p = self._p.contents
        p.metric = metric_code(metric)
        p.intermediate_graph_degree, p.graph_degree, p.nn_descent_niter = intermediate_graph_degree, graph_degree, nn_descent_niter

    def __del__(self):
        try:
            lib.cuvsCagraIndexParamsDestroy(self._p)
        except Exception:
            pass


class SearchParams:
    _ALGO = {"single_cta": 0, "multi_cta": 1, "multi_kernel": 2, "auto": 100}
    _HASH = {"hash": 0, "small": 1, "auto": 100}

    def __init__(self, *, max_queries=0, itopk_size=64, max_iterations=0, algo="auto", team_size=0, search_width=1,
                 min_iterations=0, thread_block_size=0, hashmap_mode="auto", hashmap_min_bitlen=0,
                 hashmap_max_fill_rate=0.5, num_random_samplings=1, rand_xor_mask=0x128394):
        self._p = C.POINTER(_SearchParamsC)()
        check(lib.cuvsCagraSearchParamsCreate(C.byref(self._p)))
        p = self._p.contents
        p.max_queries, p.itopk_size, p.max_iterations = max_queries, itopk_size, max_iterations
        p.algo, p.team_size, p.search_width, p.min_iterations = self._ALGO[algo], team_size, search_width, min_iterations
        p.thread_block_size, p.hashmap_mode = thread_block_size, self._HASH[hashmap_mode]
        p.hashmap_min_bitlen, p.hashmap_max_fill_rate = hashmap_min_bitlen, hashmap_max_fill_rate
        p.num_random_samplings, p.rand_xor_mask = num_random_samplings, rand_xor_mask

    def __del__(self):
        try:
            lib.cuvsCagraSearchParamsDestroy(self._p)
        except Exception:
            pass


class Index:
    def __init__(self):
        self._p = C.POINTER(index_handle)()
        check(lib.cuvsCagraIndexCreate(C.byref(self._p)))
        self.trained = False
        self._keep = None

    def __del__(self):
        try:
            if self._p:
                lib.cuvsCagraIndexDestroy(self._p)
                self._p = None
        except Exception:
            pass

    def _i64(self, fn):
        v = C.c_int64(0)
        check(fn(self._p, C.byref(v)))
        return v.value

    dim = property(lambda self: self._i64(lib.cuvsCagraIndexGetDims))
    graph_degree = property(lambda self: self._i64(lib.cuvsCagraIndexGetGraphDegree))

    def __len__(self):
        return self._i64(lib.cuvsCagraIndexGetSize)

    def set_walk_precision(self, bits, resources=None):
        """cuvs_b200 extension: 16 = walk the graph over an fp16 copy of the vectors (half the HBM gather bytes), results
        re-ranked with the fp32 rows; 32 = exact fp32 walk (default)."""
        from ..common import Resources
        res = resources or Resources()
        check(lib.cuvsB200CagraSetWalkPrecision(res.get_c_obj(), self._p, int(bits)))
        res.sync()
        return self

    @property
    def graph(self):
        m = DLManagedTensor()
        check(lib.cuvsCagraIndexGetGraph(self._p, C.byref(m)))
        return view_to_torch(m, owner=self)


@auto_sync_resources
def build(index_params, dataset, resources=None):
    ds = as_tensor(dataset)
    idx = Index()
    check(lib.cuvsCagraBuild(resources.get_c_obj(), index_params._p, DL(ds).ptr, idx._p))
    idx.trained, idx._keep = True, ds
    return idx


@auto_sync_resources
def from_graph(graph, dataset, metric="sqeuclidean", resources=None):
    """cuvsCagraIndexFromArgs: index from an existing [n, degree] uint32 graph and its dataset."""
    g, ds = as_tensor(graph), as_tensor(dataset)
    if g.dtype != torch.uint32:
        g = g.to(torch.int64).to(torch.uint32)
    idx = Index()
    check(lib.cuvsCagraIndexFromArgs(resources.get_c_obj(), C.c_int(metric_code(metric)), DL(g).ptr, DL(ds).ptr, idx._p))
    idx.trained, idx._keep = True, (g, ds)
    return idx


@auto_sync_resources
def search(search_params, index, queries, k, neighbors=None, distances=None, resources=None, filter=None):
    if not index.trained:
        raise ValueError("Index needs to be built before calling search.")
    q = as_tensor(queries)
    nq = q.shape[0]
    if neighbors is None:
        neighbors = torch.empty((nq, k), dtype=torch.uint32, device=q.device)
    if distances is None:
        distances = torch.empty((nq, k), dtype=torch.float32, device=q.device)
    f = filter.c_obj() if filter is not None else cuvsFilter(0, 0)
    check(lib.cuvsCagraSearch(resources.get_c_obj(), search_params._p, index._p, DL(q).ptr, DL(neighbors).ptr, DL(distances).ptr, f))
    return distances, neighbors


@auto_sync_resources
def save(filename, index, include_dataset=True, resources=None):
    check(lib.cuvsCagraSerialize(resources.get_c_obj(), str(filename).encode(), index._p, C.c_bool(include_dataset)))


@auto_sync_resources
def load(filename, resources=None):
    idx = Index()
    check(lib.cuvsCagraDeserialize(resources.get_c_obj(), str(filename).encode(), idx._p))
    idx.trained = True
    return idx
