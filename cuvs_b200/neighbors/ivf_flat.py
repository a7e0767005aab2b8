"""IVF-Flat — same surface as python/cuvs/cuvs/neighbors/ivf_flat/ivf_flat.pyx
(IndexParams :41, Index :153, build :207, SearchParams :265, search :295, save :397, load :434, extend :467)."""
from __future__ import annotations

import ctypes as C

import torch

from .._capi import DL, DLManagedTensor, as_tensor, check, index_handle, lib, metric_code, view_to_torch
from ..common.resources import auto_sync_resources
from .filters import no_filter


class _IndexParamsC(C.Structure):  # include/cuvs/neighbors/ivf_flat.h: struct cuvsIvfFlatIndexParams
    _fields_ = [("metric", C.c_int), ("metric_arg", C.c_float), ("add_data_on_build", C.c_bool), ("n_lists", C.c_uint32),
                ("kmeans_n_iters", C.c_uint32), ("kmeans_trainset_fraction", C.c_double), ("adaptive_centers", C.c_bool),
                ("conservative_memory_allocation", C.c_bool)]


class _SearchParamsC(C.Structure):
    _fields_ = [("n_probes", C.c_uint32)]


class IndexParams:
    def __init__(self, *, n_lists=1024, metric="sqeuclidean", metric_arg=2.0, kmeans_n_iters=20,
                 kmeans_trainset_fraction=0.5, add_data_on_build=True, adaptive_centers=False,
                 conservative_memory_allocation=False):
        self._p = C.POINTER(_IndexParamsC)()
        check(lib.cuvsIvfFlatIndexParamsCreate(C.byref(self._p)))
        p = self._p.contents
        p.metric, p.metric_arg, p.add_data_on_build, p.n_lists = metric_code(metric), metric_arg, add_data_on_build, n_lists
        p.kmeans_n_iters, p.kmeans_trainset_fraction = kmeans_n_iters, kmeans_trainset_fraction
        p.adaptive_centers, p.conservative_memory_allocation = adaptive_centers, conservative_memory_allocation

    def __del__(self):
        try:
            lib.cuvsIvfFlatIndexParamsDestroy(self._p)
        except Exception:
            pass

    n_lists = property(lambda self: self._p.contents.n_lists)
    metric = property(lambda self: self._p.contents.metric)
    kmeans_n_iters = property(lambda self: self._p.contents.kmeans_n_iters)
    add_data_on_build = property(lambda self: self._p.contents.add_data_on_build)


class SearchParams:
    def __init__(self, *, n_probes=20):
        self._p = C.POINTER(_SearchParamsC)()
        check(lib.cuvsIvfFlatSearchParamsCreate(C.byref(self._p)))
        self._p.contents.n_probes = n_probes

    def __del__(self):
        try:
            lib.cuvsIvfFlatSearchParamsDestroy(self._p)
        except Exception:
            pass

    n_probes = property(lambda self: self._p.contents.n_probes)


class Index:
    def __init__(self):
        self._p = C.POINTER(index_handle)()
        check(lib.cuvsIvfFlatIndexCreate(C.byref(self._p)))
        self.trained = False

    def __del__(self):
        try:
            if self._p:
                lib.cuvsIvfFlatIndexDestroy(self._p)
                self._p = None
        except Exception:
            pass

    def __repr__(self):
        return "Index(type=IvfFlat, metric=..., n_lists=%d, dim=%d)" % (self.n_lists, self.dim) if self.trained else "Index(type=IvfFlat)"

    def _i64(self, fn):
        v = C.c_int64(0)
        check(fn(self._p, C.byref(v)))
        return v.value

    @property
    def n_lists(self):
        return self._i64(lib.cuvsIvfFlatIndexGetNLists)

    @property
    def dim(self):
        return self._i64(lib.cuvsIvfFlatIndexGetDim)

    def __len__(self):
        return self._i64(lib.cuvsB200IvfFlatGetSize)

    @property
    def centers(self):
        m = DLManagedTensor()
        check(lib.cuvsIvfFlatIndexGetCenters(self._p, C.byref(m)))
        return view_to_torch(m, owner=self)

    @property
    def list_sizes(self):
        m = DLManagedTensor()
        check(lib.cuvsB200IvfFlatGetListSizes(self._p, C.byref(m)))
        return view_to_torch(m, owner=self)

    def list_indices(self, label):
        m = DLManagedTensor()
        check(lib.cuvsB200IvfFlatGetListIndices(self._p, C.c_uint32(label), C.byref(m)))
        return view_to_torch(m, owner=self)


@auto_sync_resources
def build(index_params, dataset, resources=None):
    ds = as_tensor(dataset)
    if ds.dtype not in (torch.float32, torch.float16, torch.int8, torch.uint8):
        raise TypeError("dtype %s not supported" % ds.dtype)
    idx = Index()
    check(lib.cuvsIvfFlatBuild(resources.get_c_obj(), index_params._p, DL(ds).ptr, idx._p))
    idx.trained = True
    return idx


@auto_sync_resources
def search(search_params, index, queries, k, neighbors=None, distances=None, resources=None, filter=None):
    if not index.trained:
        raise ValueError("Index needs to be built before calling search.")
    q = as_tensor(queries)
    nq = q.shape[0]
    if neighbors is None:
        neighbors = torch.empty((nq, k), dtype=torch.int64, device=q.device)
    if distances is None:
        distances = torch.empty((nq, k), dtype=torch.float32, device=q.device)
    f = filter or no_filter()
    check(lib.cuvsIvfFlatSearch(resources.get_c_obj(), search_params._p, index._p, DL(q).ptr, DL(neighbors).ptr,
                                DL(distances).ptr, f.c_obj()))
    return distances, neighbors


@auto_sync_resources
def extend(index, new_vectors, new_indices, resources=None):
    v = as_tensor(new_vectors)
    ids = None if new_indices is None else DL(as_tensor(new_indices).to(torch.int64))
    check(lib.cuvsIvfFlatExtend(resources.get_c_obj(), DL(v).ptr, ids.ptr if ids else None, index._p))
    return index


@auto_sync_resources
def save(filename, index, resources=None):
    check(lib.cuvsIvfFlatSerialize(resources.get_c_obj(), str(filename).encode(), index._p))


@auto_sync_resources
def load(filename, resources=None):
    idx = Index()
    check(lib.cuvsIvfFlatDeserialize(resources.get_c_obj(), str(filename).encode(), idx._p))
    idx.trained = True
    return idx


@auto_sync_resources
def set_centers(index, centers, resources=None):
    """cuvsB200IvfFlatSetCenters (extension): overwrite the coarse centres of an empty index."""
    check(lib.cuvsB200IvfFlatSetCenters(resources.get_c_obj(), index._p, DL(as_tensor(centers)).ptr))
    return index
