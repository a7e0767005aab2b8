"""GPU: recall of ivf_pq(+refine) on the bench data for a few operating points (index built once)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import gen_manifold
from cuvs_b200.neighbors import brute_force, ivf_pq, refine

n = int(os.environ.get("N", 10_000_000)); d = 128; nq = 10_000; k = 10
ds = gen_manifold(n, d, 1234); qs = gen_manifold(nq, d, 1234 + 3087)
bf = brute_force.build(ds); _, gt = brute_force.search(bf, qs, k); del bf
t0 = time.time()
index = ivf_pq.build(ivf_pq.IndexParams(n_lists=1024, pq_dim=64, pq_bits=8, kmeans_n_iters=10), ds)
torch.cuda.synchronize(); print("build s", time.time() - t0)
sizes = index.list_sizes.float(); print("list max/mean", (sizes.max() / sizes.mean()).item(), "min", sizes.min().item())
def rec(i): return (i.unsqueeze(2) == gt.unsqueeze(1)).any(dim=2).float().mean().item()
for n_probes in (16, 32, 64):
    for kc in (10, 20, 40, 64):
        sp = ivf_pq.SearchParams(n_probes=n_probes)
        torch.cuda.synchronize(); t0 = time.time()
        dd, ii = ivf_pq.search(sp, index, qs, kc)
        if kc > k:
            dd, ii = refine(ds, qs, ii, k=k)
        torch.cuda.synchronize(); dt = time.time() - t0
        print(f"n_probes={n_probes} k_cand={kc} recall@10={rec(ii[:, :k]):.4f} time_ms={dt*1e3:.2f}")
