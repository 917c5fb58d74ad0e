#!/bin/bash
# GPU call 9: the bulk-staged transform kernels and the restructured k2_fused epilogue: parity, in-situ numbers
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/c9_all_tests.txt 2>&1
tail -3 gpurun_out/c9_all_tests.txt | cut -c1-300; grep -E "^FAILED|^E  " gpurun_out/c9_all_tests.txt | head -10
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
python -c "
import json
b=[json.loads(l) for l in open('gpurun_out/c9_bench.json') if l.startswith('{')][-1]
r=b['roofline']
print('value',b['value'],'e2e',b['e2e']['value'],'frac',r['frac'],'avg_us',r['avg_launch_us'],'pc',r['preconditioner']['avg_us'],'cpu',b['cpu_baseline']['value'],b['cpu_baseline']['cores'])"
tail -2 gpurun_out/c9_bench.err
timeout 300 python bench.py --steps 2 --batch 10 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/c9_bench_first20.json 2>/dev/null
python -c "
import json
b=[json.loads(l) for l in open('gpurun_out/c9_bench_first20.json') if l.startswith('{')][-1]
r=b['roofline']
print('first 20 steps: value',b['value'],'frac',r['frac'],'avg_us',r['avg_launch_us'],'pc',r['preconditioner']['avg_us'])"
for n in 512; do timeout 300 python bench.py --grid $n --steps 2 --batch 10 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/c9_bench_512.json 2>/dev/null; done
python -c "
import json
b=[json.loads(l) for l in open('gpurun_out/c9_bench_512.json') if l.startswith('{')][-1]
r=b['roofline']
print('512^2 first 20 steps: value',b['value'],'frac',r['frac'],'avg_us',r['avg_launch_us'],'pc',r['preconditioner']['avg_us'])"
