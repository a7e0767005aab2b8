"""N > 1 on real GPUs: the list-sharded search with the exchange step inside libcuvs_c.so (ncclAllGather on the resource's
stream + merge).  Needs >= 2 GPUs (`gpurun --gpus 2`); on the one-GPU box it is skipped and the host logic is covered by
tests/test_distributed_cpu.py (gloo)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_list_sharded_search_over_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = min(torch.cuda.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29641", os.path.join(ROOT, "tests", "_dist_worker.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
