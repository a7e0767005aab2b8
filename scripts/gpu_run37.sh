#!/bin/bash
# ivf_flat workload: N = visible GPUs
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
R=gpurun_out/r37_n${N}
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --workload ivf_flat --steps 5 --warmup 3 --no-cpu > ${R}_bench.log 2>&1
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --workload ivf_flat --gpus $N --steps 5 --warmup 3 > ${R}_bench.log 2>&1
fi
python - <<PY
import json
for line in open('${R}_bench.log'):
    line=line.strip()
    if line.startswith('{'):
        j=json.loads(line)
        print(' n_gpus %d value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f frac %.3f recall %s build %s' % (j['n_gpus'], j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['config'].get('recall_at_10'), j['config'].get('index_build_s')))
    elif 'Error' in line or 'error' in line: print('  ', line[:300])
PY
tail -3 ${R}_bench.log | cut -c1-300
