"""cuVS-file-compatible (de)serialization: cuvs{IvfPq,IvfFlat,Cagra}{Serialize,Deserialize} write / read the
reference's own index file formats (ivf_pq_serialize.cuh:25-86 v4, ivf_flat_serialize.cuh:25-84 v4, cagra_serialize.cuh:30-85
v5: sequences of NumPy .npy records).  Checked from both sides with an INDEPENDENT implementation of the container — NumPy's
own np.lib.format reader / writer:
  * our files parse record by record with NumPy, carry the reference's field order / dtypes / shapes, and the list payloads
    equal the oracle's restatement of the reference layouts (interleaved PQ codes, groups-of-32 IVF-Flat rows);
  * files WRITTEN by NumPy in the reference layout load through our Deserialize and search identically."""
import io
import os

import numpy as np
import pytest
import torch

import oracle
from tests.util import clustered, uniform

pytestmark = pytest.mark.gpu


def _records(path, skip=0):
    """All NPY records of a file, in order (np.lib.format = NumPy's own reader)."""
    out = []
    with open(path, "rb") as f:
        f.read(skip)
        size = os.fstat(f.fileno()).st_size
        while f.tell() < size:
            out.append(np.lib.format.read_array(f, allow_pickle=False))
    return out


def _write(f, arr):
    np.lib.format.write_array(f, np.asarray(arr), version=(1, 0))


@pytest.mark.parametrize("pq_bits,pq_dim", [(8, 32), (5, 24)])
def test_ivf_pq_file_is_the_reference_format(tmp_path, pq_bits, pq_dim):
    from cuvs_b200.neighbors import ivf_pq as m
    ds = uniform(3000, 64, 7, 0.1, 2.0)
    qs = uniform(64, 64, 8, 0.1, 2.0)
    index = m.build(m.IndexParams(n_lists=16, pq_dim=pq_dim, pq_bits=pq_bits, kmeans_n_iters=5), torch.from_numpy(ds).cuda())
    path = str(tmp_path / "pq.bin")
    m.save(path, index)
    rec = _records(path)
    ver, size, dim, bits, pdim, cma, metric, cb, layout, n_lists = [r[()] for r in rec[:10]]
    assert (ver, size, dim, bits, pdim, metric, cb, layout, n_lists) == (4, 3000, 64, pq_bits, pq_dim, 0, 0, 1, 16)
    assert [r.dtype.str for r in rec[:10]] == ["<i4", "<i8", "<u4", "<u4", "<u4", "|u1", "<i4", "<i4", "<i4", "<u4"]
    pq_centers, centers, centers_rot, rot, sizes = rec[10:15]
    pq_len = -(-64 // pq_dim)
    assert pq_centers.shape == (pq_dim, pq_len, 1 << pq_bits) and centers.shape == (16, 72) and centers_rot.shape == (16, pq_dim * pq_len)
    assert rot.shape == (pq_dim * pq_len, 64) and sizes.dtype == np.uint32 and int(sizes.sum()) == 3000
    np.testing.assert_array_equal(pq_centers, index.pq_centers.cpu().numpy())
    np.testing.assert_array_equal(centers, index.centers_padded.cpu().numpy())
    pos = 15
    per_chunk = 128 // pq_bits
    for l in range(16):
        assert rec[pos].dtype == np.uint32 and int(rec[pos][()]) == int(sizes[l])
        pos += 1
        if sizes[l] == 0:
            continue
        data, ids = rec[pos], rec[pos + 1]
        pos += 2
        assert data.dtype == np.uint8 and data.shape == (-(-int(sizes[l]) // 32), -(-pq_dim // per_chunk), 32, 16)
        flat = index.list_data(l).cpu().numpy()  # the reference's contiguous (bit-packed row) format via the getter
        codes = oracle.unpack_pq_interleaved(data, int(sizes[l]), pq_dim, pq_bits)
        assert (oracle.pack_pq_interleaved(codes, pq_bits) == data).all()
        if pq_bits == 8:
            assert (codes == flat).all()
        assert ids.dtype == np.int64 and (ids == index.list_indices(l).cpu().numpy()).all()
    assert pos == len(rec)
    again = m.load(path)
    sp = m.SearchParams(n_probes=8)
    d0, i0 = m.search(sp, index, torch.from_numpy(qs).cuda(), 10)
    d1, i1 = m.search(sp, again, torch.from_numpy(qs).cuda(), 10)
    assert torch.equal(i0, i1) and torch.equal(d0, d1)


def test_ivf_pq_loads_a_file_written_by_numpy_in_the_reference_layout(tmp_path):
    from cuvs_b200.neighbors import ivf_pq as m
    ds = uniform(2500, 64, 9, 0.1, 2.0)
    qs = uniform(64, 64, 10, 0.1, 2.0)
    index = m.build(m.IndexParams(n_lists=8, pq_dim=32, kmeans_n_iters=5), torch.from_numpy(ds).cuda())
    sizes = index.list_sizes.cpu().numpy().astype(np.uint32)
    path = str(tmp_path / "ref_style.bin")
    with open(path, "wb") as f:
        for v, t in [(4, np.int32), (2500, np.int64), (64, np.uint32), (8, np.uint32), (32, np.uint32), (False, np.uint8), (0, np.int32),
                     (0, np.int32), (1, np.int32), (8, np.uint32)]:
            _write(f, np.array(v, dtype=t))
        _write(f, index.pq_centers.cpu().numpy())
        _write(f, index.centers_padded.cpu().numpy())
        _write(f, index.centers_rot.cpu().numpy())
        _write(f, index.rotation_matrix.cpu().numpy())
        _write(f, sizes)
        for l in range(8):
            _write(f, np.array(sizes[l], dtype=np.uint32))
            if sizes[l]:
                _write(f, oracle.pack_pq_interleaved(index.list_data(l).cpu().numpy(), 8))
                _write(f, index.list_indices(l).cpu().numpy().astype(np.int64))
    loaded = m.load(path)
    sp = m.SearchParams(n_probes=8)
    d0, i0 = m.search(sp, index, torch.from_numpy(qs).cuda(), 10)
    d1, i1 = m.search(sp, loaded, torch.from_numpy(qs).cuda(), 10)
    assert torch.equal(i0, i1) and torch.equal(d0, d1)


def test_ivf_flat_file_is_the_reference_format(tmp_path):
    from cuvs_b200.neighbors import ivf_flat as m
    ds, c = clustered(3000, 36, 3, n_centers=16)
    qs, _ = clustered(50, 36, 4, centers=c)
    index = m.build(m.IndexParams(n_lists=16, kmeans_n_iters=5), torch.from_numpy(ds).cuda())
    path = str(tmp_path / "flat.bin")
    m.save(path, index)
    with open(path, "rb") as f:
        assert f.read(4) == b"<f4\x00"
    rec = _records(path, skip=4)
    ver, size, dim, n_lists, metric, adaptive, cma = [r[()] for r in rec[:7]]
    assert (ver, size, dim, n_lists, metric) == (4, 3000, 36, 16, 0)
    assert [r.dtype.str for r in rec[:7]] == ["<i4", "<i8", "<u4", "<u4", "<i4", "|u1", "|u1"]
    centers, has_norms, norms, sizes = rec[7], rec[8], rec[9], rec[10]
    assert centers.shape == (16, 36) and bool(has_norms[()]) and sizes.dtype == np.uint32 and int(sizes.sum()) == 3000
    np.testing.assert_allclose(norms, (centers.astype(np.float64) ** 2).sum(1), rtol=1e-5)
    pos, seen = 11, []
    for l in range(16):
        cap = int(rec[pos][()])
        assert cap == -(-int(sizes[l]) // 32) * 32
        pos += 1
        if cap == 0:
            continue
        data, ids = rec[pos], rec[pos + 1]
        pos += 2
        assert data.shape == (cap, 36) and ids.shape == (cap,) and (ids[int(sizes[l]):] == -1).all()
        rows = ds[ids[:int(sizes[l])]]
        np.testing.assert_array_equal(oracle.interleave_ivf_flat(rows), data.ravel())
        seen.extend(ids[:int(sizes[l])].tolist())
    assert pos == len(rec) and sorted(seen) == list(range(3000))
    again = m.load(path)
    sp = m.SearchParams(n_probes=8)
    d0, i0 = m.search(sp, index, torch.from_numpy(qs).cuda(), 10)
    d1, i1 = m.search(sp, again, torch.from_numpy(qs).cuda(), 10)
    assert torch.equal(i0, i1) and torch.equal(d0, d1)


def test_cagra_file_is_the_reference_format(tmp_path):
    from cuvs_b200.neighbors import cagra as m
    ds, c = clustered(4000, 30, 5, n_centers=16)
    qs, _ = clustered(40, 30, 6, centers=c)
    index = m.build(m.IndexParams(graph_degree=16), torch.from_numpy(ds).cuda())
    path = str(tmp_path / "cagra.bin")
    m.save(path, index, include_dataset=True)
    with open(path, "rb") as f:
        assert f.read(4) == b"<f4\x00"
    rec = _records(path, skip=4)
    assert [int(r[()]) for r in rec[:4]] == [5, 4000, 30, 16] and [r.dtype.str for r in rec[:5]] == ["<i4", "<u4", "<u4", "<u4", "<i4"]
    graph = rec[5]
    assert graph.dtype == np.uint32 and graph.shape == (4000, 16) and (graph == index.graph.cpu().numpy()).all()
    assert int(rec[6][()]) == 1 and int(rec[7][()]) == 2 and int(rec[8][()]) == 0  # dataset follows, strided, CUDA_R_32F
    assert int(rec[9][()]) == 4000 and rec[9].dtype == np.int64 and int(rec[10][()]) == 30
    np.testing.assert_array_equal(rec[12], ds)
    assert len(rec) == 13
    again = m.load(path)
    sp = m.SearchParams(itopk_size=32)
    d0, i0 = m.search(sp, index, torch.from_numpy(qs).cuda(), 10)
    d1, i1 = m.search(sp, again, torch.from_numpy(qs).cuda(), 10)
    assert torch.equal(i0, i1) and torch.equal(d0, d1)
