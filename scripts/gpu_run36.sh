#!/bin/bash
# final round-1 run: tests, smoke, default bench + reference arm, brute force, cagra (+ ncu capture of the walk kernel)
mkdir -p gpurun_out
R=gpurun_out/r36
echo "== tests" > ${R}_tests.log
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 >> ${R}_tests.log 2>&1
echo "== smoke" > ${R}_smoke.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> ${R}_smoke.log 2>&1
echo "== default bench (ivf_pq 10M, lut f16)" > ${R}_bench.log
timeout 1500 python bench.py >> ${R}_bench.log 2>&1
echo "== brute_force" >> ${R}_bench.log
timeout 900 python bench.py --workload brute_force --steps 10 --warmup 3 --no-cpu >> ${R}_bench.log 2>&1
echo "== cagra 10M" >> ${R}_bench.log
timeout 1500 python bench.py --workload cagra --steps 10 --warmup 3 --no-cpu >> ${R}_bench.log 2>&1
echo "== reference arm" >> ${R}_bench.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 >> ${R}_bench.log 2>&1
CUVS_B200_PROFILE=1 timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:cagra_search -c 1 -o ${R}_cagra python bench.py --workload cagra --n 2000000 --steps 1 --warmup 3 --no-cpu > ${R}_ncu_cagra.log 2>&1
tail -n 4 ${R}_tests.log | cut -c1-300; tail -n 2 ${R}_smoke.log | cut -c1-300
python - <<'PY'
import json
for line in open('gpurun_out/r36_bench.log'):
    line=line.strip()
    if line.startswith('=='): print(line); continue
    if line.startswith('{'):
        j=json.loads(line)
        if j.get('impl')=='reference': print(' ref value %.2f cores %s' % (j['value'], j['cpu_baseline']['cores'])); continue
        print(' value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f frac %.3f traffic %s parity %s recall %s build %s cpu %s launches %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline']['traffic'], j['parity_spot_check'], j['config'].get('recall_at_10'), j['config'].get('index_build_s'), j.get('cpu_baseline',{}).get('value'), j['gpu_launches']))
    elif 'Error' in line or 'error' in line: print('  ', line[:300])
PY
