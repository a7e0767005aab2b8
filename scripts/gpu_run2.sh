#!/bin/bash
mkdir -p gpurun_out
echo "== tests" > gpurun_out/r2_tests.log
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 >> gpurun_out/r2_tests.log 2>&1
for epiw in 8 4; do for skip in 0 1; do
  echo "== EPIW=$epiw SKIP=$skip" >> gpurun_out/r2_bench.log
  CUVS_B200_TC_EPIW=$epiw CUVS_B200_TC_SKIP_EPI=$skip timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu >> gpurun_out/r2_bench.log 2>&1
done; done
timeout 600 python scripts/diag_flags.py > gpurun_out/r2_diag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_scan -s 3 -c 1 -o gpurun_out/r2_tc_scan python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_ncu.log 2>&1
tail -n 5 gpurun_out/r2_tests.log; grep -o '"value": [0-9.]*\|== EPIW.*\|"kernel_ms": [0-9.]*' gpurun_out/r2_bench.log; tail -n 20 gpurun_out/r2_diag.log
