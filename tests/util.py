"""Shared helpers for the parity tests (seeded inputs in the reference's own test distributions)."""
import ctypes as C

import numpy as np


def uniform(n, d, seed=1234, lo=-1.0, hi=1.0):
    """cpp/tests/neighbors/brute_force.cu:468-473: uniform(-1, 1), seed 1234."""
    return np.random.default_rng(seed).uniform(lo, hi, (n, d)).astype(np.float32)


def clustered(n, d, seed=1234, n_centers=None, sigma=0.25, centers=None):
    """mixture of Gaussians, centres ~ N(0, I), points = centre + sigma * N(0, I)."""
    rng = np.random.default_rng(seed)
    if centers is None:
        n_centers = n_centers or max(1, n // 1000)
        centers = np.random.default_rng(99).standard_normal((n_centers, d)).astype(np.float32)
    lab = rng.integers(0, centers.shape[0], n)
    return (centers[lab] + sigma * rng.standard_normal((n, d))).astype(np.float32), centers


def launches():
    from cuvs_b200._capi import lib
    lib.cuvsB200KernelLaunches.restype = C.c_longlong
    return int(lib.cuvsB200KernelLaunches())


def last_flagged():
    from cuvs_b200._capi import lib
    return int(lib.cuvsB200LastFlagged())
