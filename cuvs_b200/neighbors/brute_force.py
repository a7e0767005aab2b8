"""Exact kNN — same functions as the reference binding
(python/cuvs/cuvs/neighbors/brute_force/brute_force.pyx: Index :34, build :60, search :120, save :266, load :304)."""
from __future__ import annotations

import ctypes as C

import torch

from .._capi import DL, as_tensor, check, index_handle, lib, metric_code
from ..common.resources import auto_sync_resources
from .filters import no_filter


class Index:
    def __init__(self):
        self._p = C.POINTER(index_handle)()
        check(lib.cuvsBruteForceIndexCreate(C.byref(self._p)))
        self.trained = False
        self._keep = None  # the dataset tensor (the index holds a non-owning view of device data)

    def __del__(self):
        try:
            if self._p:
                lib.cuvsBruteForceIndexDestroy(self._p)
                self._p = None
        except Exception:
            pass

    def __repr__(self):
        return "Index(type=BruteForce)"


@auto_sync_resources
def build(dataset, metric="sqeuclidean", metric_arg=2.0, resources=None):
    ds = as_tensor(dataset)
    if ds.dtype not in (torch.float32, torch.float16, torch.int8, torch.uint8):
        raise TypeError("dtype %s not supported" % ds.dtype)
    idx = Index()
    dl = DL(ds)
    check(lib.cuvsBruteForceBuild(resources.get_c_obj(), dl.ptr, C.c_int(metric_code(metric)), C.c_float(metric_arg), idx._p))
    idx.trained = True
    idx._keep = ds
    return idx


@auto_sync_resources
def search(index, queries, k, neighbors=None, distances=None, resources=None, prefilter=None):
    if not index.trained:
        raise ValueError("Index needs to be built before calling search.")
    q = as_tensor(queries)
    if q.dtype not in (torch.float32, torch.float16, torch.int8, torch.uint8):
        raise TypeError("dtype %s not supported" % q.dtype)
    nq = q.shape[0]
    if neighbors is None:
        neighbors = torch.empty((nq, k), dtype=torch.int64, device=q.device)
    if distances is None:
        distances = torch.empty((nq, k), dtype=torch.float32, device=q.device)
    f = prefilter or no_filter()
    check(lib.cuvsBruteForceSearch(resources.get_c_obj(), index._p, DL(q).ptr, DL(neighbors).ptr, DL(distances).ptr, f.c_obj()))
    return distances, neighbors


@auto_sync_resources
def save(filename, index, include_dataset=True, resources=None):
    check(lib.cuvsBruteForceSerialize(resources.get_c_obj(), str(filename).encode(), index._p))


@auto_sync_resources
def load(filename, resources=None):
    idx = Index()
    check(lib.cuvsBruteForceDeserialize(resources.get_c_obj(), str(filename).encode(), idx._p))
    idx.trained = True
    return idx
