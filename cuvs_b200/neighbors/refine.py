"""Exact re-ranking of candidates (python/cuvs/cuvs/neighbors/refine.pyx: refine(dataset, queries, candidates, k, ...))."""
from __future__ import annotations

import ctypes as C

import torch

from .._capi import DL, as_tensor, check, lib, metric_code
from ..common.resources import auto_sync_resources


@auto_sync_resources
def refine(dataset, queries, candidates, k=None, indices=None, distances=None, metric="sqeuclidean", resources=None):
    ds, q, c = as_tensor(dataset), as_tensor(queries), as_tensor(candidates)
    nq = q.shape[0]
    if k is None:
        k = indices.shape[1] if indices is not None else distances.shape[1]
    if indices is None:
        indices = torch.empty((nq, k), dtype=torch.int64, device=q.device)
    if distances is None:
        distances = torch.empty((nq, k), dtype=torch.float32, device=q.device)
    check(lib.cuvsRefine(resources.get_c_obj(), DL(ds).ptr, DL(q).ptr, DL(c).ptr, C.c_int(metric_code(metric)), DL(indices).ptr,
                         DL(distances).ptr))
    return distances, indices
