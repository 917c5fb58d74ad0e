#!/bin/bash
# GPU call 2: values-per-thread sweep of the fast transform kernels (parity + timing), ncu full captures, config-size parity tests
mkdir -p gpurun_out
for le in 2 3 4 5; do
  echo "== BK_FFT_LOGE=$le" >> gpurun_out/c2_sweep.txt
  BK_FFT_LOGE=$le timeout 300 python -m pytest tests/test_gpu_precond.py -q -m gpu -k "dct" 2>&1 | tail -3 >> gpurun_out/c2_sweep.txt
  BK_FFT_LOGE=$le timeout 200 python tools/bench_precond.py >> gpurun_out/c2_sweep.txt 2>&1
done
cat gpurun_out/c2_sweep.txt
for le in 3 5; do
  BK_FFT_LOGE=$le timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_strided|k_contig' -s 15 -c 3 -o gpurun_out/c2_fft_e$le -f python tools/bench_precond.py 1024 > gpurun_out/c2_ncu_e$le.log 2>&1
done
timeout 1200 python -m pytest tests/test_gpu_configs.py -q -m gpu -x 2>&1 | tail -40 > gpurun_out/c2_config_tests.txt
cat gpurun_out/c2_config_tests.txt
