#!/usr/bin/env python
"""Build the IVF-PQ bench index once and sweep n_probes: QPS (device-timed, L2 flushed) and recall@10 per operating point.
usage: python scripts/sweep_probes.py N N_LISTS "24,32,48,64" [DATA_RANK [REFINE_RATIOS]]   (e.g. 32 "2,4,6" = the harder data) """
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cuvs_b200.common import Resources  # noqa: E402
from cuvs_b200.neighbors import ivf_pq  # noqa: E402

n, n_lists = int(sys.argv[1]), int(sys.argv[2])
probes = [int(x) for x in sys.argv[3].split(",")]
data_rank = int(sys.argv[4]) if len(sys.argv) > 4 else 16
refines = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [2]
wl = bench.IvfPqWorkload(n=n, n_lists=n_lists, n_probes=max(probes), data_rank=data_rank, refine_ratio=max(refines))
res = Resources()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = []
for p, rr in [(p, rr) for p in probes for rr in refines]:
    wl.sp = ivf_pq.SearchParams(n_probes=p, lut_dtype=np.float16)
    wl.refine_ratio, wl.kc = rr, wl.k * rr
    wl.cand = torch.empty((wl.nq, wl.kc), dtype=torch.int64, device="cuda")
    wl.cand_d = torch.empty((wl.nq, wl.kc), dtype=torch.float32, device="cuda")
    for _ in range(3):
        wl.step(res)
    res.sync()
    total = 0.0
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        wl.step(res)
        res.sync()
        e.record()
        e.synchronize()
        total += s.elapsed_time(e)
    if os.environ.get("CUVS_B200_PROFILE") == "1":  # ncu --profile-from-start off: one more step inside the profiler range
        torch.cuda.profiler.start()
        wl.step(res)
        res.sync()
        torch.cuda.profiler.stop()
    wl.check()
    out.append({"n": n, "n_lists": n_lists, "n_probes": p, "refine_ratio": rr, "data_rank": data_rank, "ms_per_batch": total / 5, "qps": wl.nq / (total / 5 * 1e-3), "recall_at_10": wl.recall})
    print(json.dumps(out[-1]), flush=True)
print(json.dumps({"build_s": wl.build_s}))
