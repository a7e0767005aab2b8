import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from tests.util import clustered
from cuvs_b200.neighbors import ivf_pq as m
ds, centers = clustered(6000, 128, 15, n_centers=8)
qs, _ = clustered(96, 128, 16, centers=centers)
index = m.build(m.IndexParams(n_lists=8, pq_dim=64, kmeans_n_iters=4), torch.from_numpy(ds).cuda())
os.environ["CUVS_B200_PQ_PATH"] = "tc"
os.environ["CUVS_B200_PQ_GROUP"] = sys.argv[1] if len(sys.argv) > 1 else "64"
d, i = m.search(m.SearchParams(n_probes=4, lut_dtype=np.float16), index, torch.from_numpy(qs).cuda(), 10)
torch.cuda.synchronize()
print("done", i[0].tolist())
