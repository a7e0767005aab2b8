#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python scripts/sweep_probes.py 100000000 16384 "16,24,32,48,64" > gpurun_out/r40_sweep100m.log 2>&1
tail -8 gpurun_out/r40_sweep100m.log | cut -c1-300
