#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r9_smi.log 2>&1
echo "== N=2 ivf_pq sharded" > gpurun_out/r9_bench.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 >> gpurun_out/r9_bench.log 2>&1
echo "== N=2 reference arm" >> gpurun_out/r9_bench.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 >> gpurun_out/r9_bench.log 2>&1
tail -c 3000 gpurun_out/r9_bench.log
