"""Pre-filters (mirrors python/cuvs/cuvs/neighbors/filters/filters.pyx): bitset / bitmap of uint32 words, bit = 1 keeps."""
from __future__ import annotations

import torch

from .._capi import BITMAP, BITSET, DL, NO_FILTER, as_tensor, cuvsFilter
import ctypes as C


class Prefilter:
    def __init__(self, kind: int, bits=None):
        self.kind = kind
        self._dl = None
        if bits is not None:
            t = as_tensor(bits)
            if t.dtype != torch.uint32:
                t = t.view(torch.uint32) if t.dtype == torch.int32 else t.to(torch.uint32)
            if not t.is_cuda:
                t = t.cuda()
            self._dl = DL(t.contiguous())

    def c_obj(self) -> cuvsFilter:
        if self.kind == NO_FILTER:
            return cuvsFilter(0, NO_FILTER)
        return cuvsFilter(C.addressof(self._dl.m), self.kind)


def no_filter():
    return Prefilter(NO_FILTER)


def from_bitset(bitset):
    return Prefilter(BITSET, bitset)


def from_bitmap(bitmap):
    return Prefilter(BITMAP, bitmap)
