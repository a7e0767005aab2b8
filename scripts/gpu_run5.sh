#!/bin/bash
mkdir -p gpurun_out
echo "== tests" > gpurun_out/r5_tests.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 >> gpurun_out/r5_tests.log 2>&1
echo "== brute" > gpurun_out/r5_bench.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu >> gpurun_out/r5_bench.log 2>&1
timeout 900 python scripts/recall_sweep.py > gpurun_out/r5_sweep.log 2>&1
echo "== ivf_pq tc refine4" >> gpurun_out/r5_bench.log
timeout 1200 python bench.py --workload ivf_pq --steps 5 --warmup 3 --no-cpu >> gpurun_out/r5_bench.log 2>&1
tail -n 8 gpurun_out/r5_tests.log; cat gpurun_out/r5_sweep.log; cat gpurun_out/r5_bench.log | cut -c1-1200
