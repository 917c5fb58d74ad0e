#!/bin/bash
# first GPU call of round 2: new transform kernels -- parity, timing, launch list, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt
timeout 600 python -m pytest tests/test_gpu_precond.py -q -m gpu 2>&1 | tail -40 > gpurun_out/c1_precond_tests.txt
timeout 300 python tools/bench_precond.py > gpurun_out/c1_precond_time.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_precond.py 2>&1 | tail -40 > gpurun_out/c1_all_tests.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file gpurun_out/c1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/c1_ncu_bench.log 2>&1
tail -5 gpurun_out/c1_precond_tests.txt; cat gpurun_out/c1_precond_time.txt; tail -3 gpurun_out/c1_all_tests.txt; cat gpurun_out/c1_bench.json | head -c 1500
