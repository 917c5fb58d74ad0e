#!/bin/bash
# GPU call 21: two-rank sanity run of the replicas path on the final library (short window)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c21_n2.json 2> gpurun_out/c21_n2.err
python -c "
import json
b=[json.loads(l) for l in open('gpurun_out/c21_n2.json') if l.startswith('{')][-1]
print('n_gpus',b['n_gpus'],'value',b['value'],'e2e',(b.get('e2e') or {}).get('value'),'scaling',b['scaling'],b['config'].get('parallelism'), {k:v for k,v in b['details'].items() if 'replica' in k or 'row' in k})"
tail -3 gpurun_out/c21_n2.err | cut -c1-300
