#!/bin/bash
mkdir -p gpurun_out
R=gpurun_out/r27
echo "== tests" > ${R}_tests.log
timeout 1800 python -m pytest tests -m gpu -q -x --timeout=600 >> ${R}_tests.log 2>&1
tail -n 6 ${R}_tests.log | cut -c1-300
: > ${R}_bench.log
echo "== ivf_pq 100M n_lists 16384 n_probes 64" >> ${R}_bench.log
timeout 1200 python bench.py --n 100000000 --n-lists 16384 --n-probes 64 --steps 5 --warmup 3 --no-cpu >> ${R}_bench.log 2>&1
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> ${R}_bench.log
python - <<'PY'
import json
for line in open('gpurun_out/r27_bench.log'):
    line=line.strip()
    if line.startswith('=='): print(line); continue
    if line.startswith('{'):
        j=json.loads(line)
        print(' value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f frac %.3f rows %s recall %s build %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline'].get('scanned_rows'), j['config'].get('recall_at_10'), j['config'].get('index_build_s')))
    elif 'Error' in line or 'error' in line or 'memory' in line: print('  ', line[:300])
PY
tail -5 ${R}_bench.log | cut -c1-400
