#!/bin/bash
mkdir -p gpurun_out
echo "== tests" > gpurun_out/r4_tests.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 >> gpurun_out/r4_tests.log 2>&1
echo "== brute" > gpurun_out/r4_bench.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu >> gpurun_out/r4_bench.log 2>&1
echo "== ivf_pq tc refine4" >> gpurun_out/r4_bench.log
timeout 1200 python bench.py --workload ivf_pq --steps 5 --warmup 3 --no-cpu >> gpurun_out/r4_bench.log 2>&1
echo "== ivf_pq tc refine2" >> gpurun_out/r4_bench.log
timeout 1200 python bench.py --workload ivf_pq --steps 5 --warmup 3 --no-cpu --refine-ratio 2 >> gpurun_out/r4_bench.log 2>&1
echo "== ivf_pq lut refine4 (1M)" >> gpurun_out/r4_bench.log
CUVS_B200_PQ_PATH=lut timeout 1200 python bench.py --workload ivf_pq --n 1000000 --n-lists 1024 --steps 3 --warmup 3 --no-cpu >> gpurun_out/r4_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_scan -s 3 -c 1 -o gpurun_out/r4_tc_scan python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r4_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4_pq_launches.csv python bench.py --workload ivf_pq --steps 1 --warmup 3 --no-cpu > gpurun_out/r4_ncu_pq.log 2>&1
tail -n 8 gpurun_out/r4_tests.log; cat gpurun_out/r4_bench.log | cut -c1-1500
