#!/bin/bash
# N-GPU list-sharded bench (N = number of visible GPUs)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
R=gpurun_out/r28_n${N}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > ${R}_bench.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > ${R}_ref.log 2>&1
python - <<PY
import json
for f in ['${R}_bench.log','${R}_ref.log']:
    for line in open(f):
        line=line.strip()
        if line.startswith('{'):
            j=json.loads(line)
            if j.get('impl')=='reference': print(' ref value %.2f cores %s' % (j['value'], j['cpu_baseline']['cores'])); continue
            print(' n_gpus %d value %.0f e2e %.0f ms/step %.3f kernel_ms %.3f recall %s build %s' % (j['n_gpus'], j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['config'].get('recall_at_10'), j['config'].get('index_build_s')))
        elif 'Error' in line or 'error' in line: print('  ', line[:300])
PY
