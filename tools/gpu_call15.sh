#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_complex.py tests/test_gpu_kernels.py -q -m gpu > gpurun_out/c15_tests.txt 2>&1
tail -25 gpurun_out/c15_tests.txt | cut -c1-300
timeout 200 python tools/hopf_debug.py 300 cgs2 > gpurun_out/c15_hopf.txt 2>&1; grep -E "^iter|final|FAILED" gpurun_out/c15_hopf.txt | cut -c1-250
