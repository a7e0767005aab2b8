#!/bin/bash
# first GPU bring-up: exact path, then tensor-core path, then bench + launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.log 2>&1
echo "== forced-exact tests" > gpurun_out/r1_tests_exact.log
CUVS_B200_FORCE_EXACT=1 timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "not duplicates and not candidate" >> gpurun_out/r1_tests_exact.log 2>&1
echo "== smoke" > gpurun_out/r1_smoke.log
timeout 300 python __graft_entry__.py smoke >> gpurun_out/r1_smoke.log 2>&1
echo "exit $?" >> gpurun_out/r1_smoke.log
echo "== tests" > gpurun_out/r1_tests.log
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 >> gpurun_out/r1_tests.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r1_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r1_ncu_bench.log 2>&1
tail -5 gpurun_out/r1_tests_exact.log gpurun_out/r1_smoke.log gpurun_out/r1_tests.log gpurun_out/r1_bench.log
