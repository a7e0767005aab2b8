#!/bin/bash
mkdir -p gpurun_out
CUVS_B200_PROFILE=1 timeout 1700 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:tc_scan -c 2 -o gpurun_out/r42_ivfpq100m_tc_scan python scripts/sweep_probes.py 100000000 16384 "48" > gpurun_out/r42_ncu.log 2>&1
tail -4 gpurun_out/r42_ncu.log | cut -c1-300
