timeout 300 python scripts/pq_debug.py 1 > gpurun_out/pq_debug3.log 2>&1; grep -c CASE_OK gpurun_out/pq_debug3.log
if [ $(grep -c CASE_OK gpurun_out/pq_debug3.log) -ge 7 ]; then bash scripts/gpu_stage.sh pq_tests c2; fi
