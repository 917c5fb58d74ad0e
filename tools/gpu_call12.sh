#!/bin/bash
# GPU call 12: contiguous transform kernels without row staging (natural-order array with its own padding): parity + in-situ time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_precond.py tests/test_gpu_kernels.py tests/test_gpu_palc.py -x -q -m gpu > gpurun_out/c12_tests.txt 2>&1
tail -3 gpurun_out/c12_tests.txt | cut -c1-300; grep -E "^FAILED|^E  " gpurun_out/c12_tests.txt | head
for le in 2 4 5; do BK_FFT_LOGE=$le timeout 300 python -m pytest tests/test_gpu_precond.py -q -m gpu -k dct 2>&1 | tail -1; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
python -c "
import json
b=[json.loads(l) for l in open('gpurun_out/c12_bench.json') if l.startswith('{')][-1]
r=b['roofline']
print('value',b['value'],'frac',r['frac'],'avg_us',r['avg_launch_us'],'pc',r['preconditioner']['avg_us'], b['details']['rejected_steps'])"
tail -2 gpurun_out/c12_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -k regex:'k_strided|k_contig' -s 30 -c 30 --csv --log-file gpurun_out/c12_fft_warm.csv python tools/bench_precond.py 1024 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/c12_fft_warm.csv
