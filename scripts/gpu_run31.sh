#!/bin/bash
mkdir -p gpurun_out
R=gpurun_out/r31
echo "== tests" > ${R}_tests.log
timeout 1800 python -m pytest tests -m gpu -q --timeout=600 >> ${R}_tests.log 2>&1
tail -n 25 ${R}_tests.log | cut -c1-300
