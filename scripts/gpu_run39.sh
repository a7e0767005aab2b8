#!/bin/bash
mkdir -p gpurun_out
R=gpurun_out/r39
echo "== tests" > ${R}_tests.log
timeout 1800 python -m pytest tests -m gpu -q --timeout=600 >> ${R}_tests.log 2>&1
tail -n 5 ${R}_tests.log | cut -c1-300
: > ${R}_bench.log
echo "== ivf_pq 10M" >> ${R}_bench.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu >> ${R}_bench.log 2>&1
echo "== brute" >> ${R}_bench.log
timeout 600 python bench.py --workload brute_force --steps 10 --warmup 3 --no-cpu >> ${R}_bench.log 2>&1
python - <<'PY'
import json
for line in open('gpurun_out/r39_bench.log'):
    line=line.strip()
    if line.startswith('=='): print(line); continue
    if line.startswith('{'):
        j=json.loads(line)
        print(' value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f frac %.3f parity %s recall %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['parity_spot_check'], j['config'].get('recall_at_10')))
    elif 'Error' in line or 'error' in line: print('  ', line[:300])
PY
