"""Resources handle (mirrors python/cuvs/cuvs/common/resources.pyx:18-75).

Wraps a ``cuvsResources_t``.  By default the handle is bound to torch's current CUDA stream so
that library work is ordered with the caller's torch work; ``sync()`` is ``cuvsStreamSync``.
"""
from __future__ import annotations

import ctypes as C
import functools

import torch

from .._capi import check, lib


class Resources:
    def __init__(self, stream=None, device=None):
        if device is not None:
            torch.cuda.set_device(device)
        self._h = C.c_size_t(0)
        check(lib.cuvsResourcesCreate(C.byref(self._h)))
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        elif hasattr(stream, "cuda_stream"):
            stream = stream.cuda_stream
        check(lib.cuvsStreamSet(self._h, C.c_void_p(stream)))

    def get_c_obj(self):
        return self._h

    def sync(self):
        check(lib.cuvsStreamSync(self._h))

    @property
    def device_id(self) -> int:
        d = C.c_int(0)
        check(lib.cuvsDeviceIdGet(self._h, C.byref(d)))
        return d.value

    def __del__(self):
        try:
            if self._h.value:
                lib.cuvsResourcesDestroy(self._h)
                self._h = C.c_size_t(0)
        except Exception:
            pass


def auto_sync_resources(f):
    """If no ``resources=`` is passed, create one for the call and sync before returning
    (python/cuvs/cuvs/common/resources.pyx:78-110)."""

    @functools.wraps(f)
    def wrapper(*args, resources=None, **kwargs):
        sync = resources is None
        res = resources or Resources()
        out = f(*args, resources=res, **kwargs)
        if sync:
            res.sync()
        return out

    return wrapper
