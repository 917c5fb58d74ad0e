#!/bin/bash
# GPU call 13: complexified contexts (complex shifts, J', Hopf minimally augmented Newton) + regression of the kernel tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_complex.py -q -m gpu > gpurun_out/c13_complex.txt 2>&1
tail -25 gpurun_out/c13_complex.txt | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/c13_kernels.txt 2>&1
tail -3 gpurun_out/c13_kernels.txt | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
