#!/bin/bash
mkdir -p gpurun_out
R=gpurun_out/r19
: > ${R}_bench.log
for f in 1 769 1025 513; do
  echo "== small L2-resident index, dbg flags $f (stages = f>>8)" >> ${R}_bench.log
  CUVS_B200_TC_SKIP_EPI=$f timeout 600 python bench.py --n 300000 --n-lists 32 --n-probes 32 --steps 10 --warmup 3 --no-cpu >> ${R}_bench.log 2>&1
done
python - <<'PY'
import json
for line in open('gpurun_out/r19_bench.log'):
    line=line.strip()
    if line.startswith('=='): print(line); continue
    if line.startswith('{'):
        j=json.loads(line)
        print(' value %.0f ms/step %.3f kernel_ms %.3f frac %.3f rows %s recall %s' % (j['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline'].get('scanned_rows'), j['config'].get('recall_at_10')))
    elif 'Error' in line or 'error' in line: print('  ', line[:300])
PY
