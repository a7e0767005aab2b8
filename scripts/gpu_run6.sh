#!/bin/bash
mkdir -p gpurun_out
echo "== tests" > gpurun_out/r6_tests.log
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 >> gpurun_out/r6_tests.log 2>&1
echo "== smoke" >> gpurun_out/r6_tests.log
timeout 300 python __graft_entry__.py smoke >> gpurun_out/r6_tests.log 2>&1
echo "== ivf_pq default" > gpurun_out/r6_bench.log
timeout 1200 python bench.py --steps 5 --warmup 3 >> gpurun_out/r6_bench.log 2>&1
tail -n 40 gpurun_out/r6_tests.log; cat gpurun_out/r6_bench.log | cut -c1-2500
