#!/bin/bash
mkdir -p gpurun_out
R=gpurun_out/r23
CUVS_B200_TC_SKIP_EPI=1 CUVS_B200_PROFILE=1 timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:tc_scan -c 2 -o ${R}_small_skip python bench.py --n 300000 --n-lists 32 --n-probes 32 --steps 1 --warmup 3 --no-cpu > ${R}_ncu.log 2>&1
tail -3 ${R}_ncu.log
