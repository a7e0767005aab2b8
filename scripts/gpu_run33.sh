#!/bin/bash
mkdir -p gpurun_out
R=gpurun_out/r33
echo "== tests" > ${R}_tests.log
timeout 1800 python -m pytest tests -m gpu -q --timeout=600 >> ${R}_tests.log 2>&1
tail -n 12 ${R}_tests.log | cut -c1-300
: > ${R}_bench.log
echo "== cagra 10M fp16 walk" >> ${R}_bench.log
timeout 1500 python bench.py --workload cagra --steps 10 --warmup 3 --no-cpu >> ${R}_bench.log 2>&1
echo "== cagra 10M fp32 walk" >> ${R}_bench.log
timeout 1500 python bench.py --workload cagra --steps 10 --warmup 3 --no-cpu --walk-bits 32 >> ${R}_bench.log 2>&1
python - <<'PY'
import json
for line in open('gpurun_out/r33_bench.log'):
    line=line.strip()
    if line.startswith('=='): print(line); continue
    if line.startswith('{'):
        j=json.loads(line)
        print(' value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f frac %.3f recall %s build %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['config'].get('recall_at_10'), j['config'].get('index_build_s')))
    elif 'Error' in line or 'error' in line: print('  ', line[:300])
PY
