"""Diagnostics: why does the brute-force certificate flag queries on the bench data? (GPU)"""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import gen_clustered
from cuvs_b200._capi import DL, check, lib
from cuvs_b200.common import Resources
from cuvs_b200.neighbors import brute_force

n, d, nq, k = 1_000_000, 128, 10_000, 10
g = torch.Generator(device="cuda"); g.manual_seed(99)
centers = torch.randn((n // 1000, d), generator=g, device="cuda")
ds = gen_clustered(n, d, 1234, centers)
qs = gen_clustered(nq, d, 1234 + 3087, centers)
index = brute_force.build(ds)
res = Resources()
for KC in (16, 32):
    pos = torch.zeros(nq, KC, dtype=torch.uint32, device="cuda")
    sc = torch.zeros(nq, KC, dtype=torch.float32, device="cuda")
    check(lib.cuvsB200BruteForceCandidates(res.get_c_obj(), index._p, DL(qs).ptr, DL(pos).ptr, DL(sc).ptr))
    p = pos.to(torch.int64)
    qn = (qs.double() ** 2).sum(1)
    approx = qn[:, None] + 2 * sc.double()
    exact = ((qs[:, None, :].double() - ds[p].double()) ** 2).sum(-1)
    err = (approx - exact).abs()
    xn_max = (ds.double() ** 2).sum(1).max()
    eps = (qn + xn_max) / 8192.0
    ek = exact.sort(dim=1).values[:, k - 1]
    aw = approx.max(dim=1).values
    margin = aw - ek
    print(f"KC={KC}: max abs err {err.max().item():.3e}  rel {(err / (qn[:, None] + xn_max)).max().item():.3e}  eps range "
          f"{eps.min().item():.4f}..{eps.max().item():.4f}  xn_max {xn_max.item():.1f}")
    print(f"   margin (A_worst - E_k) quantiles:", [round(margin.quantile(q).item(), 4) for q in (0.0, 0.001, 0.01, 0.1, 0.5)])
    print(f"   flagged (margin <= eps): {(margin <= eps).sum().item()}   with eps/4: {(margin <= eps / 4).sum().item()}")
    bad = (margin <= eps).nonzero().flatten()[:5]
    for b in bad.tolist():
        print("   q", b, "E sorted", [round(v, 4) for v in exact[b].sort().values.tolist()], "eps", round(eps[b].item(), 4))
