#!/bin/bash
mkdir -p gpurun_out
echo "== tests" > gpurun_out/r3_tests.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x >> gpurun_out/r3_tests.log 2>&1
echo "== rest" >> gpurun_out/r3_tests.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 >> gpurun_out/r3_tests.log 2>&1
for skip in 0 1; do
  echo "== EPIW=8 SKIP=$skip" >> gpurun_out/r3_bench.log
  CUVS_B200_TC_SKIP_EPI=$skip timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu >> gpurun_out/r3_bench.log 2>&1
done
tail -n 30 gpurun_out/r3_tests.log; grep -o '"value": [0-9.]*\|== EPIW.*\|"kernel_ms": [0-9.]*' gpurun_out/r3_bench.log
