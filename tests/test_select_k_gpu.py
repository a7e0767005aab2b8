"""GPU parity: cuvsSelectK / cuvsKnnMergeParts vs the oracle (ties -> smaller position, padding)."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _select(v, k, select_min=True, idx=None, out_dtype=torch.int64):
    from cuvs_b200._capi import DL, check, lib
    from cuvs_b200.common import Resources
    res = Resources()
    tv = torch.from_numpy(v).cuda()
    ov = torch.empty(v.shape[0], k, device="cuda")
    oi = torch.empty(v.shape[0], k, dtype=out_dtype, device="cuda")
    ti = None if idx is None else DL(torch.from_numpy(idx).cuda())
    check(lib.cuvsSelectK(res.get_c_obj(), DL(tv).ptr, ti.ptr if ti else None, DL(ov).ptr, DL(oi).ptr,
                          C.c_bool(select_min), C.c_bool(True)))
    res.sync()
    return ov.cpu().numpy(), oi.cpu().numpy().astype(np.int64)


@pytest.mark.parametrize("batch,length,k", [(7, 1, 1), (33, 100, 10), (10, 640, 10), (4, 5000, 64), (3, 70000, 100),
                                            (2, 300, 300), (5, 1024, 1024), (2, 9, 20), (1, 1 << 20, 16)])
@pytest.mark.parametrize("select_min", [True, False])
def test_select_k_matches_oracle(batch, length, k, select_min):
    rng = np.random.default_rng(batch * 1000 + length + k)
    v = rng.standard_normal((batch, length)).astype(np.float32)
    v[:, ::7] = np.round(v[:, ::7], 1)  # plenty of exact ties
    ov, oi = _select(v, k, select_min)
    rv, ri = oracle.select_k(v, k, select_min)
    np.testing.assert_array_equal(oi, ri)
    np.testing.assert_array_equal(ov, rv)


@pytest.mark.parametrize("length,k", [(2048, 48), (4096, 1), (4097, 256), (16384, 48), (20000, 100), (32768, 64), (16384, 300)])
@pytest.mark.parametrize("kind", ["normal", "narrow", "all_equal", "three_values", "with_inf", "sorted_asc", "sorted_desc", "planted"])
def test_select_k_medium_rows_register_path(length, k, kind):
    """Rows of 2k..32k elements with k <= 256 take the register-resident kernel (linear binning over the live span, crowded
    buckets re-binned): same answers as the oracle incl. the tie rule, for value distributions that stress the binning."""
    rng = np.random.default_rng(length + k)
    batch = 5
    if kind == "normal":
        v = rng.standard_normal((batch, length)).astype(np.float32)
        v[:, ::5] = np.round(v[:, ::5], 1)
    elif kind == "narrow":          # coarse-search distances: one exponent, tiny spread, many exact ties
        v = (100.0 + 1e-4 * rng.integers(0, 50, (batch, length))).astype(np.float32)
    elif kind == "all_equal":
        v = np.full((batch, length), 3.25, np.float32)
    elif kind == "three_values":
        v = rng.choice(np.array([-1.0, 0.0, 7.5], np.float32), (batch, length))
    elif kind == "planted":   # one outlier per warp inside the sample, nothing else near: fewer than k keys under the sampled bound
        v = (100.0 + rng.standard_normal((batch, length))).astype(np.float32)
        v[:, 0:512:32] = -1000.0 - np.arange(16, dtype=np.float32)
        v[:, 7] = 1e6
    elif kind in ("sorted_asc", "sorted_desc"):   # the kernel's sampled bound (first elements of every thread) misses: full-span retry
        v = np.sort(rng.standard_normal((batch, length)).astype(np.float32), axis=1)
        if kind == "sorted_desc":
            v = np.ascontiguousarray(v[:, ::-1])
    else:
        v = rng.standard_normal((batch, length)).astype(np.float32)
        v[:, ::3] = np.inf
        v[:, 1::97] = -np.inf
    for select_min in (True, False):
        ov, oi = _select(v, k, select_min)
        rv, ri = oracle.select_k(v, k, select_min)
        np.testing.assert_array_equal(oi, ri)
        np.testing.assert_array_equal(ov, rv)
    idx = rng.permutation(batch * length).astype(np.int64).reshape(batch, length)
    ov, oi = _select(v, min(k, 64), True, idx)
    rv, ri = oracle.select_k(v, min(k, 64), True, idx)
    np.testing.assert_array_equal(oi, ri)


def test_select_k_with_payload_and_special_values():
    v = np.array([[np.inf, -np.inf, 0.0, -0.0, 5.0, 5.0, 1e-38, 3.4e38]], np.float32)
    idx = (np.arange(8, dtype=np.int64) * 10 + 3)[None, :]
    ov, oi = _select(v, 4, True, idx)
    assert oi[0].tolist() == [13, 23, 33, 63]  # -inf, 0.0 == -0.0 (position order), 1e-38
    ov, oi = _select(v, 3, False, idx)
    assert oi[0].tolist() == [3, 73, 43]


def test_knn_merge_parts():
    from cuvs_b200._capi import DL, check, lib
    from cuvs_b200.common import Resources
    rng = np.random.default_rng(5)
    n_parts, n_rows, k = 8, 100, 10
    keys = np.sort(rng.random((n_parts, n_rows, k)).astype(np.float32), axis=2)
    vals = rng.integers(0, 1000, (n_parts, n_rows, k)).astype(np.int64)
    trans = (np.arange(n_parts) * 1000).astype(np.int64)
    res = Resources()
    ik = torch.from_numpy(keys.reshape(n_parts * n_rows, k)).cuda()
    iv = torch.from_numpy(vals.reshape(n_parts * n_rows, k)).cuda()
    ok = torch.empty(n_rows, k, device="cuda")
    ov = torch.empty(n_rows, k, dtype=torch.int64, device="cuda")
    tr = (C.c_int64 * n_parts)(*trans.tolist())
    check(lib.cuvsKnnMergeParts(res.get_c_obj(), DL(ik).ptr, DL(iv).ptr, DL(ok).ptr, DL(ov).ptr, C.c_int64(n_parts), tr,
                                C.c_bool(True)))
    res.sync()
    allk = keys.transpose(1, 0, 2).reshape(n_rows, n_parts * k)
    allv = (vals + trans[:, None, None]).transpose(1, 0, 2).reshape(n_rows, n_parts * k)
    rk, ri = oracle.select_k(allk, k, True, allv)
    np.testing.assert_array_equal(ok.cpu().numpy(), rk)
    np.testing.assert_array_equal(ov.cpu().numpy(), ri)
