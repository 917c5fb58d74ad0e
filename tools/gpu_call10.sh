#!/bin/bash
# GPU call 10 (--gpus 8): replicas mode of bench.py on 8 ranks over NCCL, as the driver launches it
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c10_bench_n8.json 2> gpurun_out/c10_bench_n8.err
grep "^{" gpurun_out/c10_bench_n8.json | cut -c1-200; grep -o '"details".*"clocks"' gpurun_out/c10_bench_n8.json | cut -c1-1800; grep -o '"e2e".*' gpurun_out/c10_bench_n8.json | cut -c1-400; tail -3 gpurun_out/c10_bench_n8.err | cut -c1-300
