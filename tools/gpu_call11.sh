#!/bin/bash
# GPU call 11: what the driver runs at round end -- GPU tests, smoke(), bench N = 1 (default flags) -- on the final commit
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c11_all_tests.txt 2>&1
tail -3 gpurun_out/c11_all_tests.txt | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/c11_bench_default.json 2> gpurun_out/c11_bench_default.err
python -c "
import json
b=[json.loads(l) for l in open('gpurun_out/c11_bench_default.json') if l.startswith('{')][-1]
r=b['roofline']
print('value',b['value'],'steps',b['steps'],'warmup',b['warmup'],'e2e',b['e2e']['value'],b['e2e']['corrector_work'],'frac',r['frac'],'pc',r['preconditioner']['avg_us'],'cpu',b['cpu_baseline']['value'],b['cpu_baseline']['cores'], b['details']['rejected_steps'])"
tail -2 gpurun_out/c11_bench_default.err
