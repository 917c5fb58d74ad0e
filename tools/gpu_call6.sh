#!/bin/bash
# GPU call 6 (--gpus 2): replicas mode of bench.py on 2 ranks over NCCL
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c6_bench_n2.json 2> gpurun_out/c6_bench_n2.err
tail -c 2500 gpurun_out/c6_bench_n2.json; tail -5 gpurun_out/c6_bench_n2.err
