#!/bin/bash
mkdir -p gpurun_out
echo "== tests" > gpurun_out/r8_tests.log
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 >> gpurun_out/r8_tests.log 2>&1
echo "== cagra 1M" > gpurun_out/r8_bench.log
timeout 1200 python bench.py --workload cagra --n 1000000 --steps 5 --warmup 3 --no-cpu >> gpurun_out/r8_bench.log 2>&1
echo "== cagra 10M" >> gpurun_out/r8_bench.log
timeout 1800 python bench.py --workload cagra --steps 5 --warmup 3 --no-cpu >> gpurun_out/r8_bench.log 2>&1
tail -n 12 gpurun_out/r8_tests.log | cut -c1-300
python - <<'PY'
import json
for line in open('gpurun_out/r8_bench.log'):
    line=line.strip()
    if line.startswith('=='): print(line); continue
    if line.startswith('{'):
        j=json.loads(line)
        print(' value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f frac %.3f parity %s recall %s build %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['parity_spot_check'], j['config'].get('recall_at_10'), j['config'].get('index_build_s')))
    elif 'Error' in line or 'error' in line: print('  ', line[:300])
PY
