#!/bin/bash
# GPU call 8: compact evidence run -- tests, ncu captures reduced to CSV on the box (reports deleted: 64 MiB limit), launch list, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/c8_all_tests.txt 2>&1
tail -3 gpurun_out/c8_all_tests.txt | cut -c1-300
timeout 600 ncu --set full --clock-control none -k regex:'k2_apply|k_sh_apply|k_cgl_apply|k_potrap|k_reduce|k_axpby|k_scale|k_gen|k_tail|k_lincomb' -c 44 -o gpurun_out/c8_tour -f python tools/kernel_tour.py > gpurun_out/c8_tour.log 2>&1
python tools/ncu_summary.py gpurun_out/c8_tour.ncu-rep gpurun_out/c8_ncu_tour.csv > /dev/null 2>&1; rm -f gpurun_out/c8_tour.ncu-rep
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'k2_fused|k2_update' -s 2400 -c 4 -o gpurun_out/c8_pair_late -f python bench.py --steps 3 --batch 10 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/c8_pair_late.log 2>&1
python tools/ncu_summary.py gpurun_out/c8_pair_late.ncu-rep gpurun_out/c8_ncu_pair_late.csv > /dev/null 2>&1; rm -f gpurun_out/c8_pair_late.ncu-rep
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'k_strided|k_contig' -s 300 -c 3 -o gpurun_out/c8_fft -f python bench.py --steps 1 --batch 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/c8_fft.log 2>&1
python tools/ncu_summary.py gpurun_out/c8_fft.ncu-rep gpurun_out/c8_ncu_fft.csv > /dev/null 2>&1; rm -f gpurun_out/c8_fft.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1200 --csv --log-file gpurun_out/c8_launches.csv python bench.py --steps 1 --batch 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/c8_launches.csv 2>/dev/null | head -9
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err
python -c "
import json
b=[json.loads(l) for l in open('gpurun_out/c8_bench.json') if l.startswith('{')][-1]
print('value',b['value'],'e2e',b['e2e']['value'],'frac',b['roofline']['frac'],'pc',b['roofline']['preconditioner']['avg_us'],'cpu',b['cpu_baseline']['value'],b['cpu_baseline']['cores'])"
tail -2 gpurun_out/c8_bench.err
ls -la gpurun_out; du -sh gpurun_out
