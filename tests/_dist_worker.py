"""Worker of tests/test_distributed_nccl.py: run under torchrun with 2+ ranks, one GPU each.  List-sharded IVF-Flat through
cuvs_b200.distributed with the library's own NCCL communicator (csrc/comm.cu); every rank checks the merged result against
the oracle's exact kNN (n_probes = n_lists makes the sharded search exact) and that all ranks hold identical results."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import oracle  # noqa: E402
from tests.util import clustered  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from cuvs_b200.common import Resources
    from cuvs_b200.distributed import build_sharded_ivf_flat
    from cuvs_b200.neighbors import ivf_flat
    ds, centers = clustered(40000, 64, 31, n_centers=32)
    qs, _ = clustered(400, 64, 32, centers=centers)
    res = Resources()
    dsg = torch.from_numpy(ds).cuda()
    ids = torch.arange(len(ds), dtype=torch.int64, device="cuda")
    chunks = ((dsg[s:s + 10000], ids[s:s + 10000]) for s in range(0, len(ds), 10000))
    sh = build_sharded_ivf_flat(ivf_flat.IndexParams(n_lists=64, kmeans_n_iters=10), dsg[::4].contiguous(), chunks, resources=res)
    assert sh.comm is not None, "the library's NCCL communicator must carry the exchange step"
    owned = int(sh.local.list_sizes.sum().item())
    tot = torch.tensor([owned], device="cuda")
    dist.all_reduce(tot)
    assert int(tot.item()) == len(ds) and 0 < owned < len(ds), (owned, int(tot.item()))
    d, i = sh.search(ivf_flat.SearchParams(n_probes=64), torch.from_numpy(qs).cuda(), 10, resources=res)
    res.sync()
    gd, gi = oracle.knn(ds, qs, 10)
    assert oracle.recall_with_ties(i.cpu().numpy(), d.cpu().numpy(), gi, gd, eps=1e-4) >= 0.999
    assert (i.cpu().numpy() == gi).mean() >= 0.995
    # identical on every rank
    ref = i.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, i)
    dist.barrier()
    if rank == 0:
        print(f"DIST_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
