#!/bin/bash
# GPU call 3: tables in shared memory (E sweep again), config-size tests, scout probe, new bench N=1
mkdir -p gpurun_out
for le in 2 3 4 5; do
  echo "== BK_FFT_LOGE=$le" >> gpurun_out/c3_sweep.txt
  BK_FFT_LOGE=$le timeout 300 python -m pytest tests/test_gpu_precond.py -q -m gpu -k "dct" 2>&1 | tail -3 >> gpurun_out/c3_sweep.txt
  BK_FFT_LOGE=$le timeout 200 python tools/bench_precond.py 512 1024 2048 >> gpurun_out/c3_sweep.txt 2>&1
done
cat gpurun_out/c3_sweep.txt
timeout 1500 python -m pytest tests/test_gpu_configs.py -q -m gpu 2>&1 | tail -60 > gpurun_out/c3_config_tests.txt
cat gpurun_out/c3_config_tests.txt | tail -40
timeout 900 python tools/scout_probe.py 1.0 > gpurun_out/c3_scout_probe.txt 2>&1
cat gpurun_out/c3_scout_probe.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
tail -c 3000 gpurun_out/c3_bench.json; tail -5 gpurun_out/c3_bench.err
