#!/bin/bash
mkdir -p gpurun_out
R=gpurun_out/r44
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 > ${R}_tests.log 2>&1
tail -n 3 ${R}_tests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${R}_smoke.log 2>&1; tail -n 1 ${R}_smoke.log
timeout 600 python bench.py > ${R}_bench.log 2>&1; tail -n 1 ${R}_bench.log | cut -c1-1500
