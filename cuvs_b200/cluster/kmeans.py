"""k-means — same surface as python/cuvs/cuvs/cluster/kmeans/kmeans.pyx (KMeansParams, fit, predict, cluster_cost)."""
from __future__ import annotations

import ctypes as C

import torch

from .._capi import DL, as_tensor, check, lib, metric_code
from ..common.resources import auto_sync_resources


class _ParamsC(C.Structure):  # include/cuvs/cluster/kmeans.h: struct cuvsKMeansParams
    _fields_ = [("metric", C.c_int), ("n_clusters", C.c_int), ("init", C.c_int), ("max_iter", C.c_int), ("tol", C.c_double),
                ("n_init", C.c_int), ("oversampling_factor", C.c_double), ("batch_samples", C.c_int), ("batch_centroids", C.c_int),
                ("inertia_check", C.c_bool), ("hierarchical", C.c_bool), ("hierarchical_n_iters", C.c_int),
                ("streaming_batch_size", C.c_int64), ("init_size", C.c_int64)]


class KMeansParams:
    _INIT = {"k-means++": 0, "kmeans++": 0, "random": 1, "array": 2}

    def __init__(self, *, metric="sqeuclidean", n_clusters=8, init_method="k-means++", max_iter=300, tol=1e-4, n_init=1,
                 hierarchical=False, hierarchical_n_iters=20):
        self._p = C.POINTER(_ParamsC)()
        check(lib.cuvsKMeansParamsCreate(C.byref(self._p)))
        p = self._p.contents
        p.metric, p.n_clusters, p.init, p.max_iter, p.tol, p.n_init = metric_code(metric), n_clusters, self._INIT[init_method], max_iter, tol, n_init
        p.hierarchical, p.hierarchical_n_iters = hierarchical, hierarchical_n_iters

    def __del__(self):
        try:
            lib.cuvsKMeansParamsDestroy(self._p)
        except Exception:
            pass

    n_clusters = property(lambda self: self._p.contents.n_clusters)


@auto_sync_resources
def fit(params, X, centroids=None, sample_weights=None, resources=None):
    x = as_tensor(X)
    if centroids is None:
        centroids = torch.empty((params.n_clusters, x.shape[1]), dtype=torch.float32, device="cuda")
    inertia, n_iter = C.c_double(0), C.c_int(0)
    w = DL(as_tensor(sample_weights)) if sample_weights is not None else None
    check(lib.cuvsKMeansFit(resources.get_c_obj(), params._p, DL(x).ptr, w.ptr if w else None, DL(centroids).ptr, C.byref(inertia),
                            C.byref(n_iter)))
    return centroids, inertia.value, n_iter.value


@auto_sync_resources
def predict(params, X, centroids, sample_weights=None, labels=None, normalize_weight=True, resources=None):
    x = as_tensor(X)
    if labels is None:
        labels = torch.empty(x.shape[0], dtype=torch.int32, device="cuda")
    inertia = C.c_double(0)
    w = DL(as_tensor(sample_weights)) if sample_weights is not None else None
    check(lib.cuvsKMeansPredict(resources.get_c_obj(), params._p, DL(x).ptr, w.ptr if w else None, DL(as_tensor(centroids)).ptr,
                                DL(labels).ptr, C.c_bool(normalize_weight), C.byref(inertia)))
    return labels, inertia.value


@auto_sync_resources
def cluster_cost(X, centroids, resources=None):
    cost = C.c_double(0)
    check(lib.cuvsKMeansClusterCost(resources.get_c_obj(), DL(as_tensor(X)).ptr, DL(as_tensor(centroids)).ptr, C.byref(cost)))
    return cost.value
