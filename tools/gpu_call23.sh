#!/bin/bash
# GPU call 23: codim-2 curves on the device (continuation_fold on SH2d hexagons, continuation_hopf on cGL2d)
mkdir -p gpurun_out
timeout 130 python -m pytest tests/test_gpu_codim2_curves.py -q -m gpu -p no:cacheprovider --durations=5 > gpurun_out/c23_codim2_tests.txt 2>&1
echo "pytest rc $?" >> gpurun_out/c23_codim2_tests.txt
tail -60 gpurun_out/c23_codim2_tests.txt | cut -c1-400
