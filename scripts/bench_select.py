"""Micro-benchmark of cuvsSelectK on the coarse-search shape (rows of n_lists distances, k = n_probes)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from cuvs_b200._capi import DL, check, lib
from cuvs_b200.common import Resources

nq, n, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
res = Resources()
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.randn((nq, 128), device="cuda", generator=g)
c = torch.randn((n, 128), device="cuda", generator=g)
v = torch.cdist(x, c).pow(2).contiguous()
ov = torch.empty(nq, k, device="cuda"); oi = torch.empty(nq, k, dtype=torch.int64, device="cuda")
def run():
    check(lib.cuvsSelectK(res.get_c_obj(), DL(v).ptr, None, DL(ov).ptr, DL(oi).ptr, C.c_bool(True), C.c_bool(True)))
for _ in range(3): run()
res.sync(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize(); res.sync()
print("select_k", nq, n, k, "ms", e0.elapsed_time(e1) / 10)
tv, ti = torch.topk(v, k, dim=1, largest=False)
print("ids equal torch.topk:", float((ti == oi).float().mean()))
