"""ctypes binding of the in-tree libcuvs_c.so (the C-ABI drop-in built by cuvs_b200/build.py).

This is the same boundary the reference's Cython layer binds (python/cuvs/cuvs/common/c_api.pxd,
cydlpack.pyx): DLManagedTensor* in, cuvsError_t out, error text via cuvsGetLastErrorText().
There is deliberately NO fallback: if the library is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CUVS_B200_LIB") or os.path.join(_HERE, "lib", "libcuvs_c.so")  # (override: A/B builds of the same sources)


class CuvsError(RuntimeError):
    """Raised when a C call returns CUVS_ERROR (mirrors cuvs.common.exceptions.CuvsException)."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m cuvs_b200.build` "
            "(cuvs_b200 has no CPU or PyTorch fallback path)")
    return C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


lib = _load()
lib.cuvsGetLastErrorText.restype = C.c_char_p


# ----------------------------------------------------------------------------- DLPack (v0.8 ABI)
class DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", DLDevice), ("ndim", C.c_int32), ("dtype", DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class DLManagedTensor(C.Structure):
    pass


DLManagedTensor._fields_ = [("dl_tensor", DLTensor), ("manager_ctx", C.c_void_p),
                            ("deleter", C.CFUNCTYPE(None, C.POINTER(DLManagedTensor)))]

kDLCPU, kDLCUDA, kDLCUDAHost = 1, 2, 3
kDLInt, kDLUInt, kDLFloat = 0, 1, 2

_TORCH_DT = {
    torch.float32: (kDLFloat, 32), torch.float16: (kDLFloat, 16), torch.float64: (kDLFloat, 64),
    torch.int64: (kDLInt, 64), torch.int32: (kDLInt, 32), torch.int8: (kDLInt, 8), torch.uint8: (kDLUInt, 8),
    torch.uint32: (kDLUInt, 32),
}


class index_handle(C.Structure):
    """cuvs*Index structs: { uintptr_t addr; DLDataType dtype; }"""
    _fields_ = [("addr", C.c_size_t), ("dtype", DLDataType)]


class cuvsFilter(C.Structure):
    _fields_ = [("addr", C.c_size_t), ("type", C.c_int)]


NO_FILTER, BITSET, BITMAP = 0, 1, 2


def as_tensor(a, device=None) -> torch.Tensor:
    """Accept torch tensors / numpy arrays / anything with __cuda_array_interface__."""
    if isinstance(a, torch.Tensor):
        return a
    if isinstance(a, np.ndarray):
        return torch.from_numpy(a)
    if hasattr(a, "__cuda_array_interface__"):
        return torch.as_tensor(a, device="cuda")
    return torch.as_tensor(a)


class DL:
    """Owns a DLManagedTensor describing a torch tensor (no copy); keeps the tensor alive."""

    def __init__(self, t: torch.Tensor, force_strides: bool = False):
        self.t = t
        code, bits = _TORCH_DT[t.dtype]
        nd = t.dim()
        self._shape = (C.c_int64 * max(nd, 1))(*t.shape)
        self._strides = (C.c_int64 * max(nd, 1))(*t.stride())
        m = DLManagedTensor()
        m.dl_tensor.data = t.data_ptr()
        if t.is_cuda:
            m.dl_tensor.device = DLDevice(kDLCUDA, t.device.index or 0)
        elif t.is_pinned():
            m.dl_tensor.device = DLDevice(kDLCUDAHost, 0)
        else:
            m.dl_tensor.device = DLDevice(kDLCPU, 0)
        m.dl_tensor.ndim = nd
        m.dl_tensor.dtype = DLDataType(code, bits, 1)
        m.dl_tensor.shape = C.cast(self._shape, C.POINTER(C.c_int64))
        if t.is_contiguous() and not force_strides:
            m.dl_tensor.strides = None
        else:
            m.dl_tensor.strides = C.cast(self._strides, C.POINTER(C.c_int64))
        m.dl_tensor.byte_offset = 0
        m.manager_ctx = None
        self.m = m

    @property
    def ptr(self):
        return C.byref(self.m)


def empty_dl():
    return DLManagedTensor()


def view_to_torch(m: DLManagedTensor, owner=None) -> torch.Tensor:
    """Wrap a library-filled *view* (index getters) as a torch tensor via __cuda_array_interface__."""
    t = m.dl_tensor
    shape = tuple(t.shape[i] for i in range(t.ndim))
    code, bits = t.dtype.code, t.dtype.bits
    typestr = {(kDLFloat, 32): "<f4", (kDLFloat, 16): "<f2", (kDLInt, 64): "<i8", (kDLUInt, 32): "<u4",
               (kDLInt, 32): "<i4", (kDLUInt, 8): "|u1", (kDLInt, 8): "|i1"}[(code, bits)]

    class _V:
        pass

    v = _V()
    v.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(t.data or 0), False), "version": 3,
                                  "strides": None}
    v._owner = owner
    if any(s == 0 for s in shape):
        out = torch.empty(shape, dtype={"<f4": torch.float32, "<i8": torch.int64, "<u4": torch.uint32,
                                        "<i4": torch.int32, "|u1": torch.uint8, "|i1": torch.int8,
                                        "<f2": torch.float16}[typestr], device="cuda")
    else:
        out = torch.as_tensor(v, device=f"cuda:{t.device.device_id}")
    out._cuvs_owner = owner
    if m.deleter:
        m.deleter(C.byref(m))
    return out


def check(status: int):
    if status != 1:  # CUVS_SUCCESS
        txt = lib.cuvsGetLastErrorText()
        raise CuvsError(txt.decode() if txt else "cuVS call failed")


METRICS = {
    "l2": 1, "sqeuclidean": 0, "euclidean": 1, "l1": 3, "cityblock": 3, "inner_product": 6, "chebyshev": 7,
    "canberra": 8, "cosine": 2, "lp": 9, "correlation": 10, "jaccard": 11, "hellinger": 12, "braycurtis": 14,
    "jensenshannon": 15, "hamming": 16, "kl_divergence": 17, "minkowski": 9, "russellrao": 18, "dice": 19,
    "bitwise_hamming": 20, "l2_unexpanded": 4, "l2_sqrt_unexpanded": 5,
}
METRIC_NAMES = {0: "sqeuclidean", 1: "euclidean", 2: "cosine", 6: "inner_product", 4: "l2_unexpanded",
                5: "l2_sqrt_unexpanded"}


def metric_code(m) -> int:
    if isinstance(m, str):
        if m not in METRICS:
            raise ValueError(f"metric {m!r} is not supported")
        return METRICS[m]
    return int(m)
