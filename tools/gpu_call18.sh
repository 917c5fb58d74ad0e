#!/bin/bash
# GPU call 18: round-end verification on the final commit -- all GPU tests, smoke(), ncu --set full of the final transform kernels, bench N = 1 (default flags)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c18_all_tests.txt 2>&1
tail -3 gpurun_out/c18_all_tests.txt | cut -c1-300; grep -E "^FAILED|^E  " gpurun_out/c18_all_tests.txt | head -8 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none -k regex:'k_strided|k_contig' -s 30 -c 3 -o gpurun_out/c18_fft -f python tools/bench_precond.py 1024 > gpurun_out/c18_fft.log 2>&1
python tools/ncu_summary.py gpurun_out/c18_fft.ncu-rep gpurun_out/c18_ncu_fft.csv > /dev/null 2>&1; rm -f gpurun_out/c18_fft.ncu-rep
timeout 900 python bench.py > gpurun_out/c18_bench_default.json 2> gpurun_out/c18_bench_default.err
python -c "
import json
b=[json.loads(l) for l in open('gpurun_out/c18_bench_default.json') if l.startswith('{')][-1]
r=b['roofline']
print('value',b['value'],'steps',b['steps'],'warmup',b['warmup'],'e2e',b['e2e']['value'],b['e2e']['corrector_work'],'frac',r['frac'],'pc',r['preconditioner']['avg_us'],'cpu',b['cpu_baseline']['value'],b['cpu_baseline']['cores'], b['details']['rejected_steps'])"
tail -2 gpurun_out/c18_bench_default.err
du -sh gpurun_out
