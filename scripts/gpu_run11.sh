#!/bin/bash
# new scan kernel (half-norm K extension, 4-deep TMEM ring), rank-binned item order: correctness first, then speed
mkdir -p gpurun_out
R=gpurun_out/r11
echo "== tests" > ${R}_tests.log
timeout 1800 python -m pytest tests -m gpu -q -x --timeout=600 >> ${R}_tests.log 2>&1
tail -n 15 ${R}_tests.log | cut -c1-300
echo "== bench" > ${R}_bench.log
for wl in ivf_pq brute_force; do
  echo "== $wl" >> ${R}_bench.log
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu >> ${R}_bench.log 2>&1
done
echo "== ivf_pq skip-epilogue" >> ${R}_bench.log
CUVS_B200_TC_SKIP_EPI=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu >> ${R}_bench.log 2>&1
echo "== brute skip-epilogue" >> ${R}_bench.log
CUVS_B200_TC_SKIP_EPI=1 timeout 900 python bench.py --workload brute_force --steps 5 --warmup 3 --no-cpu >> ${R}_bench.log 2>&1
python - <<'PY'
import json
for line in open('gpurun_out/r11_bench.log'):
    line=line.strip()
    if line.startswith('=='): print(line); continue
    if line.startswith('{'):
        j=json.loads(line)
        print(' value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f frac %.3f parity %s recall %s build %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['parity_spot_check'], j['config'].get('recall_at_10'), j['config'].get('index_build_s')))
    elif 'Error' in line or 'error' in line: print('  ', line[:300])
PY
