"""CPU / gloo, world_size 2: the host logic of the list-sharded search (ownership rule, all-gather layout, merge) with the
oracle standing in for the per-shard scan and the merge kernel.  Sharded result must equal the unsharded oracle search."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cuvs_b200.distributed import ShardedIvfFlat, owner_of_list
    rng = np.random.default_rng(0)
    n, d, n_lists, n_probes, k, nq = 3000, 16, 12, 5, 7, 40
    ds = rng.standard_normal((n, d)).astype(np.float32)
    qs = rng.standard_normal((nq, d)).astype(np.float32)
    centers = ds[:n_lists].copy()
    labels, _ = oracle.kmeans_assign(ds, centers)
    order = np.argsort(labels, kind="stable")
    sizes = np.bincount(labels, minlength=n_lists)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    # this rank's shard: same centres, only the lists it owns
    own = owner_of_list(np.arange(n_lists), world) == rank
    keep = own[labels[order]]
    sh_sizes = np.where(own, sizes, 0)
    sh_offs = np.concatenate([[0], np.cumsum(sh_sizes)])
    shard = (centers, sh_offs, ds[order][keep], order[keep].astype(np.int64))

    def local_search(local, sp, queries, kk):
        c, o, x, ids = local
        dd, ii = oracle.ivf_flat_search(c, o, x, ids, queries.numpy(), sp, kk)
        return torch.from_numpy(dd), torch.from_numpy(ii)

    def merge(keys, vals, n_parts, kk, select_min):
        nq_ = keys.shape[0] // n_parts
        kk_ = keys.view(n_parts, nq_, kk).permute(1, 0, 2).reshape(nq_, n_parts * kk).numpy()
        vv_ = vals.view(n_parts, nq_, kk).permute(1, 0, 2).reshape(nq_, n_parts * kk).numpy()
        ov, oi = oracle.select_k(kk_, kk, select_min, vv_)
        return torch.from_numpy(ov), torch.from_numpy(oi)

    idx = ShardedIvfFlat(shard, local_search=local_search, merge=merge)
    dd, ii = idx.search(n_probes, torch.from_numpy(qs), k)
    if rank == 0:
        rd, ri = oracle.ivf_flat_search(centers, offs, ds[order], order.astype(np.int64), qs, n_probes, k)
        out["ok"] = bool((ii.numpy() == ri).all() and np.allclose(dd.numpy(), rd))
        out["world"] = idx.world
    dist.barrier()
    dist.destroy_process_group()


def test_list_sharded_search_equals_unsharded():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["world"] == 2 and out["ok"]


def test_owner_rule_balances_lists():
    from cuvs_b200.distributed import owner_of_list
    ids = torch.arange(1000)
    counts = torch.bincount(owner_of_list(ids, 8), minlength=8)
    assert counts.max() - counts.min() <= 1
