"""List-sharded IVF-Flat search across the GPUs of one box: one process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch) for the single exchange step.

Reference behaviour being replaced: cpp/src/neighbors/mg/snmg.cuh:248-375 shards the dataset by ROWS, trains an independent
index per shard, and gathers partial results to a root rank with ncclSend/ncclRecv + knn_merge_parts, returning through
the host.  Here (BASELINE.json north_star) the index is sharded by IVF LIST:
  * the coarse centres are replicated (trained on rank 0, broadcast), so every rank computes identical probe lists;
  * list l lives on rank  owner(l) = l % world  (lists are balanced by k-means, so a modulo placement balances bytes);
  * a search scans, on each rank, only the probed lists that rank owns (the library drops probes of empty lists on the
    device) and yields a partial top-k [nq, k] with GLOBAL ids;
  * ONE all-gather of the packed partials (nq*k*12 bytes per rank) + a k-way merge (cuvsKnnMergeParts) on every rank.
There is no other data-path collective; the result is identical on all ranks.

The host logic (ownership rule, gather layout, merge call) is backend-agnostic: tests/test_distributed_cpu.py drives it
with gloo on CPU tensors through `local_search=` / `merge=` hooks that use the oracle.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def owner_of_list(list_id, world: int):
    """Rank that stores IVF list `list_id` (works on ints and tensors)."""
    return list_id % world


def _default_merge(keys: torch.Tensor, vals: torch.Tensor, n_parts: int, k: int, select_min: bool):
    """[n_parts * nq, k] part-major partials -> [nq, k] via the library's cuvsKnnMergeParts."""
    import ctypes as C

    from ._capi import DL, check, lib
    from .common.resources import Resources
    nq = keys.shape[0] // n_parts
    ok = torch.empty((nq, k), dtype=torch.float32, device=keys.device)
    ov = torch.empty((nq, k), dtype=torch.int64, device=keys.device)
    res = Resources()
    h = [DL(keys), DL(vals), DL(ok), DL(ov)]
    check(lib.cuvsKnnMergeParts(res.get_c_obj(), h[0].ptr, h[1].ptr, h[2].ptr, h[3].ptr, C.c_int64(n_parts), None, C.c_bool(select_min)))
    res.sync()
    return ok, ov


class Comm:
    """The library's own NCCL communicator (csrc/comm.cu): one per process/GPU.  Bootstrapped over an existing
    torch.distributed group — rank 0 makes the ncclUniqueId inside the library, the 128 bytes are broadcast, every rank calls
    cuvsB200CommCreate.  After that the exchange step of a sharded search runs entirely inside libcuvs_c.so on the
    resource's stream (pack -> ONE ncclAllGather -> k-way merge), with no host synchronisation."""

    def __init__(self, resources, group=None):
        import ctypes as C

        from ._capi import check, lib
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        ident = (C.c_ubyte * 128)()
        if self.rank == 0:
            check(lib.cuvsB200NcclUniqueId(ident))
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.tensor(list(ident), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0, group=group)
        ident = (C.c_ubyte * 128)(*t.cpu().tolist())
        self._p = C.c_void_p()
        check(lib.cuvsB200CommCreate(resources.get_c_obj(), ident, C.c_int(self.rank), C.c_int(self.world), C.byref(self._p)))

    def allgather_merge(self, resources, d, i, out_d, out_i, select_min=True):
        import ctypes as C

        from ._capi import DL, check, lib
        h = [DL(d), DL(i), DL(out_d), DL(out_i)]  # (held across the call: a DL owns the shape array its DLTensor points at)
        check(lib.cuvsB200AllGatherMergeTopK(resources.get_c_obj(), self._p, h[0].ptr, h[1].ptr, h[2].ptr, h[3].ptr, C.c_bool(select_min)))

    def __del__(self):
        try:
            from ._capi import lib
            if self._p:
                lib.cuvsB200CommDestroy(self._p)
                self._p = None
        except Exception:
            pass


class ShardedIvfFlat:
    """A list-sharded IVF index: `local` holds the lists this rank owns (all other lists are empty).

    comm: a `Comm` (the library's NCCL communicator) — the exchange step then runs inside libcuvs_c.so on the resource's
    stream.  Without one (CPU tests over gloo; `local_search=` / `merge=` hooks) the same layout is exchanged with
    torch.distributed collectives."""

    def __init__(self, local_index, group=None, select_min: bool = True,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None, comm: Optional["Comm"] = None):
        self.local = local_index
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.select_min = select_min
        self._local_search = local_search
        self._merge = merge or _default_merge
        self.comm = comm
        self._out = None

    def search(self, search_params, queries, k, resources=None):
        """Every rank passes the same `queries`; returns (distances [nq,k], global ids [nq,k]) — identical on all ranks."""
        if self._local_search is not None:
            d, i = self._local_search(self.local, search_params, queries, k)
        else:
            from .neighbors import ivf_flat
            d, i = ivf_flat.search(search_params, self.local, queries, k, resources=resources)
        if self.world == 1:
            return d, i
        nq = d.shape[0]
        if self.comm is not None and resources is not None:
            if self._out is None or self._out[0].shape != d.shape:
                self._out = (torch.empty_like(d), torch.empty_like(i))
            self.comm.allgather_merge(resources, d, i, self._out[0], self._out[1], self.select_min)
            return self._out
        keys = torch.empty((self.world * nq, k), dtype=d.dtype, device=d.device)
        vals = torch.empty((self.world * nq, k), dtype=i.dtype, device=i.device)
        # the one exchange step: all-gather of the partial top-k (part-major layout = what knn_merge_parts takes)
        dist.all_gather_into_tensor(keys, d.contiguous(), group=self.group)
        dist.all_gather_into_tensor(vals, i.contiguous(), group=self.group)
        return self._merge(keys, vals, self.world, k, self.select_min)


def build_sharded_ivf_flat(index_params, train_rows: torch.Tensor, chunks, group=None, resources=None):
    """Build this rank's shard.

    train_rows : [n_train, dim] rows used to train the coarse centres (any rank's copy works; rank 0's centres win).
    chunks     : iterable of (rows [m, dim] device tensor, global_ids [m] int64 device tensor); every rank iterates over the
                 SAME chunks (synthetic data is regenerated per rank from the seed; a real loader would read a shared file)
                 and keeps the rows whose list it owns.  No inter-rank data movement.
    """
    from .cluster import kmeans
    from .neighbors import ivf_flat
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n_lists = index_params.n_lists
    p = ivf_flat.IndexParams(n_lists=n_lists, metric=index_params.metric, kmeans_n_iters=index_params.kmeans_n_iters,
                             kmeans_trainset_fraction=1.0, add_data_on_build=False)
    index = ivf_flat.build(p, train_rows, resources=resources)
    centers = index.centers.clone()
    if world > 1:
        dist.broadcast(centers, src=0, group=group)  # bit-identical partition rule on every rank
        ivf_flat.set_centers(index, centers, resources=resources)
    kp = kmeans.KMeansParams(n_clusters=n_lists)
    for rows, ids in chunks:
        labels, _ = kmeans.predict(kp, rows, centers, resources=resources)
        mine = owner_of_list(labels.to(torch.int64), world) == rank
        if bool(mine.any()):
            ivf_flat.extend(index, rows[mine].contiguous(), ids[mine].contiguous(), resources=resources)
    comm = Comm(resources, group) if (world > 1 and resources is not None and dist.get_backend(group) == "nccl") else None
    return ShardedIvfFlat(index, group=group, comm=comm)
