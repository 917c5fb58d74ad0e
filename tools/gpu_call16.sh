#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_complex.py -q -m gpu > gpurun_out/c16_tests.txt 2>&1
tail -40 gpurun_out/c16_tests.txt | cut -c1-300
