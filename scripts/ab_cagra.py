#!/usr/bin/env python
"""A/B of the CAGRA walk's L2 prefetch on ONE index: CUVS_B200_CAGRA_PREFETCH = 0 (off) / 1 (per-lane line prefetch) / 2 (bulk
prefetch, UBLKPF) — the library reads the variable per search.  usage: python scripts/ab_cagra.py [N] [WALK_BITS]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cuvs_b200.common import Resources  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 32
wl = bench.CagraWorkload(n=n, walk_bits=bits)
res = Resources()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode in ["0", "1", "2", "0", "1", "2"]:
    os.environ["CUVS_B200_CAGRA_PREFETCH"] = mode
    for _ in range(3):
        wl.step(res)
    res.sync()
    total = 0.0
    for _ in range(8):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        wl.step(res)
        res.sync()
        e.record()
        e.synchronize()
        total += s.elapsed_time(e)
    wl.check()
    print(json.dumps({"prefetch": mode, "n": n, "walk_bits": bits, "ms_per_batch": total / 8, "qps": wl.nq / (total / 8 * 1e-3),
                      "recall_at_10": wl.recall}), flush=True)
