#!/bin/bash
mkdir -p gpurun_out
CUVS_B200_TC_PREFETCH=1 timeout 1500 python scripts/sweep_probes.py 100000000 16384 "48" > gpurun_out/r41_sweep100m_prefetch.log 2>&1
tail -3 gpurun_out/r41_sweep100m_prefetch.log | cut -c1-300
