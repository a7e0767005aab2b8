"""IVF-PQ — same surface as python/cuvs/cuvs/neighbors/ivf_pq/ivf_pq.pyx
(IndexParams :40, Index :239, build :477, build_precomputed :543, SearchParams :667, search :745, save :843,
load :880, extend :913, transform :987)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .._capi import DL, DLManagedTensor, as_tensor, check, index_handle, lib, metric_code, view_to_torch
from ..common.resources import auto_sync_resources

CUDA_R_32F, CUDA_R_16F, CUDA_R_8U, CUDA_R_8I = 0, 2, 8, 3


class _IndexParamsC(C.Structure):  # include/cuvs/neighbors/ivf_pq.h: struct cuvsIvfPqIndexParams
    _fields_ = [("metric", C.c_int), ("metric_arg", C.c_float), ("add_data_on_build", C.c_bool), ("n_lists", C.c_uint32),
                ("kmeans_n_iters", C.c_uint32), ("kmeans_trainset_fraction", C.c_double), ("pq_bits", C.c_uint32),
                ("pq_dim", C.c_uint32), ("codebook_kind", C.c_int), ("force_random_rotation", C.c_bool),
                ("conservative_memory_allocation", C.c_bool), ("max_train_points_per_pq_code", C.c_uint32),
                ("codes_layout", C.c_int)]


class _SearchParamsC(C.Structure):
    _fields_ = [("n_probes", C.c_uint32), ("lut_dtype", C.c_int), ("internal_distance_dtype", C.c_int),
                ("coarse_search_dtype", C.c_int), ("max_internal_batch_size", C.c_uint32),
                ("preferred_shmem_carveout", C.c_double)]


def _cuda_dtype(dt):
    dt = np.dtype(dt)
    if dt == np.float32:
        return CUDA_R_32F
    if dt == np.float16:
        return CUDA_R_16F
    if dt == np.uint8:
        return CUDA_R_8U
    if dt == np.int8:
        return CUDA_R_8I
    raise ValueError("unsupported dtype %s" % dt)


class IndexParams:
    def __init__(self, *, n_lists=1024, metric="sqeuclidean", metric_arg=2.0, kmeans_n_iters=20,
                 kmeans_trainset_fraction=0.5, pq_bits=8, pq_dim=0, codebook_kind="subspace", force_random_rotation=False,
                 add_data_on_build=True, conservative_memory_allocation=False, max_train_points_per_pq_code=256,
                 codes_layout="interleaved"):
        self._p = C.POINTER(_IndexParamsC)()
        check(lib.cuvsIvfPqIndexParamsCreate(C.byref(self._p)))
        p = self._p.contents
        p.metric, p.metric_arg, p.add_data_on_build, p.n_lists = metric_code(metric), metric_arg, add_data_on_build, n_lists
        p.kmeans_n_iters, p.kmeans_trainset_fraction = kmeans_n_iters, kmeans_trainset_fraction
        p.pq_bits, p.pq_dim = pq_bits, pq_dim
        p.codebook_kind = {"subspace": 0, "cluster": 1}[codebook_kind]
        p.force_random_rotation = force_random_rotation
        p.conservative_memory_allocation = conservative_memory_allocation
        p.max_train_points_per_pq_code = max_train_points_per_pq_code
        p.codes_layout = {"flat": 0, "interleaved": 1}[codes_layout]

    def __del__(self):
        try:
            lib.cuvsIvfPqIndexParamsDestroy(self._p)
        except Exception:
            pass


class SearchParams:
    def __init__(self, *, n_probes=20, lut_dtype=np.float32, internal_distance_dtype=np.float32,
                 coarse_search_dtype=np.float32, max_internal_batch_size=4096):
        self._p = C.POINTER(_SearchParamsC)()
        check(lib.cuvsIvfPqSearchParamsCreate(C.byref(self._p)))
        p = self._p.contents
        p.n_probes = n_probes
        p.lut_dtype = _cuda_dtype(lut_dtype)
        p.internal_distance_dtype = _cuda_dtype(internal_distance_dtype)
        p.coarse_search_dtype = _cuda_dtype(coarse_search_dtype)
        p.max_internal_batch_size = max_internal_batch_size

    def __del__(self):
        try:
            lib.cuvsIvfPqSearchParamsDestroy(self._p)
        except Exception:
            pass


class Index:
    def __init__(self):
        self._p = C.POINTER(index_handle)()
        check(lib.cuvsIvfPqIndexCreate(C.byref(self._p)))
        self.trained = False

    def __del__(self):
        try:
            if self._p:
                lib.cuvsIvfPqIndexDestroy(self._p)
                self._p = None
        except Exception:
            pass

    def _i64(self, fn):
        v = C.c_int64(0)
        check(fn(self._p, C.byref(v)))
        return v.value

    n_lists = property(lambda self: self._i64(lib.cuvsIvfPqIndexGetNLists))
    dim = property(lambda self: self._i64(lib.cuvsIvfPqIndexGetDim))
    pq_dim = property(lambda self: self._i64(lib.cuvsIvfPqIndexGetPqDim))
    pq_len = property(lambda self: self._i64(lib.cuvsIvfPqIndexGetPqLen))
    pq_bits = property(lambda self: self._i64(lib.cuvsIvfPqIndexGetPqBits))

    def __len__(self):
        return self._i64(lib.cuvsIvfPqIndexGetSize)

    def _info(self):
        path, nbytes = C.c_int(0), C.c_int64(0)
        check(lib.cuvsB200IvfPqIndexInfo(self._p, C.byref(path), C.byref(nbytes)))
        return path.value, nbytes.value

    streamed = property(lambda self: (self._info()[0] & 2) != 0)  # holds the code stream (code-streaming scan, scan_pq.cu)
    has_decoded_rows = property(lambda self: (self._info()[0] & 1) != 0)  # decoded bf16 rows (small-index cache / decoded-row scan)
    device_bytes = property(lambda self: self._info()[1])        # bytes of device memory the index holds

    def _view(self, fn, *args):
        m = DLManagedTensor()
        check(fn(self._p, *args, C.byref(m)))
        return view_to_torch(m, owner=self)

    centers = property(lambda self: self._view(lib.cuvsIvfPqIndexGetCenters))
    centers_padded = property(lambda self: self._view(lib.cuvsIvfPqIndexGetCentersPadded))
    pq_centers = property(lambda self: self._view(lib.cuvsIvfPqIndexGetPqCenters))
    centers_rot = property(lambda self: self._view(lib.cuvsIvfPqIndexGetCentersRot))
    rotation_matrix = property(lambda self: self._view(lib.cuvsIvfPqIndexGetRotationMatrix))
    list_sizes = property(lambda self: self._view(lib.cuvsIvfPqIndexGetListSizes))

    def list_indices(self, label):
        return self._view(lib.cuvsIvfPqIndexGetListIndices, C.c_uint32(label))

    def list_data(self, label, n_rows=0, offset=0, resources=None):
        from ..common.resources import Resources
        res = resources or Resources()
        size = int(self.list_sizes[label].item())
        n_rows = n_rows or size - offset
        ld = (self.pq_dim * self.pq_bits + 7) // 8
        out = torch.empty((n_rows, ld), dtype=torch.uint8, device="cuda")
        check(lib.cuvsIvfPqIndexUnpackContiguousListData(res.get_c_obj(), self._p, DL(out).ptr, C.c_uint32(label), C.c_uint32(offset)))
        res.sync()
        return out


@auto_sync_resources
def build(index_params, dataset, resources=None):
    ds = as_tensor(dataset)
    if ds.dtype not in (torch.float32, torch.float16, torch.int8, torch.uint8):
        raise TypeError("dtype %s not supported" % ds.dtype)
    idx = Index()
    check(lib.cuvsIvfPqBuild(resources.get_c_obj(), index_params._p, DL(ds).ptr, idx._p))
    idx.trained = True
    return idx


@auto_sync_resources
def build_precomputed(index_params, dim, pq_centers, centers, centers_rot=None, rotation_matrix=None, resources=None):
    idx = Index()
    dls = [DL(as_tensor(pq_centers)), DL(as_tensor(centers)),
           DL(as_tensor(centers_rot)) if centers_rot is not None else None,
           DL(as_tensor(rotation_matrix)) if rotation_matrix is not None else None]
    check(lib.cuvsIvfPqBuildPrecomputed(resources.get_c_obj(), index_params._p, C.c_uint32(dim), dls[0].ptr, dls[1].ptr,
                                        dls[2].ptr if dls[2] else None, dls[3].ptr if dls[3] else None, idx._p))
    idx.trained = True
    return idx


@auto_sync_resources
def search(search_params, index, queries, k, neighbors=None, distances=None, resources=None):
    if not index.trained:
        raise ValueError("Index needs to be built before calling search.")
    q = as_tensor(queries)
    nq = q.shape[0]
    if neighbors is None:
        neighbors = torch.empty((nq, k), dtype=torch.int64, device=q.device)
    if distances is None:
        distances = torch.empty((nq, k), dtype=torch.float32, device=q.device)
    check(lib.cuvsIvfPqSearch(resources.get_c_obj(), search_params._p, index._p, DL(q).ptr, DL(neighbors).ptr, DL(distances).ptr))
    return distances, neighbors


@auto_sync_resources
def extend(index, new_vectors, new_indices, resources=None):
    v = as_tensor(new_vectors)
    ids = None if new_indices is None else DL(as_tensor(new_indices).to(torch.int64))
    check(lib.cuvsIvfPqExtend(resources.get_c_obj(), DL(v).ptr, ids.ptr if ids else None, index._p))
    return index


@auto_sync_resources
def transform(index, input_dataset, output_labels=None, output_dataset=None, resources=None):
    x = as_tensor(input_dataset)
    n = x.shape[0]
    ld = (index.pq_dim * index.pq_bits + 7) // 8
    if output_labels is None:
        output_labels = torch.empty(n, dtype=torch.uint32, device=x.device)
    if output_dataset is None:
        output_dataset = torch.empty((n, ld), dtype=torch.uint8, device=x.device)
    check(lib.cuvsIvfPqTransform(resources.get_c_obj(), index._p, DL(x).ptr, DL(output_labels).ptr, DL(output_dataset).ptr))
    return output_labels, output_dataset


@auto_sync_resources
def save(filename, index, resources=None):
    check(lib.cuvsIvfPqSerialize(resources.get_c_obj(), str(filename).encode(), index._p))


@auto_sync_resources
def load(filename, resources=None):
    idx = Index()
    check(lib.cuvsIvfPqDeserialize(resources.get_c_obj(), str(filename).encode(), idx._p))
    idx.trained = True
    return idx
