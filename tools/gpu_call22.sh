#!/bin/bash
# GPU call 22 (the round's last GPU minutes): bk_palc_run on the B200 -- parity tests of the native loop against the plugin-surface
# loop, then the two loops timed on the benchmark's own branch at several grid sizes (tools/native_loop_check.py).  No torch import.
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_gpu_native_loop.py -q -m gpu -p no:cacheprovider > gpurun_out/c22_native_tests.txt 2>&1
echo "pytest rc $?" >> gpurun_out/c22_native_tests.txt
tail -15 gpurun_out/c22_native_tests.txt
timeout 100 python tools/native_loop_check.py --grid 1024 512 256 --steps 30 --out gpurun_out/c22_native_loop_check.json > gpurun_out/c22_native_check.log 2>&1
echo "check rc $?" >> gpurun_out/c22_native_check.log
tail -8 gpurun_out/c22_native_check.log | cut -c1-1500
