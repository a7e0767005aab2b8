import os, sys, subprocess, numpy as np, torch
sys.path.insert(0, os.getcwd())
import oracle
from tests.util import clustered
which = sys.argv[1]
if which == "inv":
    from cuvs_b200.neighbors import ivf_pq as m
    for dim, pqd in ((128, 64), (64, 32)):
        ds, centers = clustered(40000, dim, 15, n_centers=64)
        qs, _ = clustered(512, dim, 16, centers=centers)
        index = m.build(m.IndexParams(n_lists=64, pq_dim=pqd, kmeans_n_iters=10), torch.from_numpy(ds).cuda())
        os.environ["CUVS_B200_PQ_PATH"] = "lut"
        dl, il = m.search(m.SearchParams(n_probes=8, lut_dtype=np.float16), index, torch.from_numpy(qs).cuda(), 10)
        os.environ["CUVS_B200_PQ_PATH"] = "tc"
        il = il.cpu().numpy()
        if dim == 64: break
        for (g, nob) in [("32", "0"), ("64", "0"), ("64", "0"), ("64", "0"), ("128", "0"), ("128", "0"), ("128", "0")]:
            os.environ["CUVS_B200_PQ_GROUP"] = g
            os.environ["CUVS_B200_PQ_MODE"] = nob
            d, i = m.search(m.SearchParams(n_probes=8, lut_dtype=np.float16), index, torch.from_numpy(qs).cuda(), 10)
            i = i.cpu().numpy()
            inter = np.mean([len(np.intersect1d(a, b)) / 10.0 for a, b in zip(i, il)])
            print(f"dim={dim} mode={nob} group={g}: id-set overlap with the LUT kernel {inter:.4f}", flush=True)
elif which == "bf":
    from cuvs_b200.neighbors import brute_force
    rng = np.random.default_rng(1)
    for (n, d, nq, k) in [(8096, 32, 128, 8), (8192, 32, 128, 8), (8096, 64, 128, 8), (1000, 32, 16, 8)]:
        ds = rng.uniform(0.1, 2.0, (n, d)).astype(np.float32); qs = rng.uniform(0.1, 2.0, (nq, d)).astype(np.float32)
        try:
            index = brute_force.build(torch.from_numpy(ds).cuda())
            dd, ii = brute_force.search(index, torch.from_numpy(qs).cuda(), k)
            torch.cuda.synchronize()
            rd, ri = oracle.knn(ds, qs, k)
            print((n, d, nq, k), "ok", (ii.cpu().numpy() == ri).mean(), flush=True)
        except Exception as e:
            print((n, d, nq, k), "FAILED", repr(e)[:300], flush=True)
            break
elif which == "ref":
    for exe in ("ref_core_c_api", "ref_c_drivers"):
        r = subprocess.run([os.path.join("oracle", "_ref", exe)], capture_output=True, text=True)
        print(exe, "rc", r.returncode, r.stdout[-1500:], r.stderr[-1500:], flush=True)
