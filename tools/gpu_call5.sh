#!/bin/bash
# GPU call 5: all GPU tests, racecheck / memcheck summaries, small ncu full capture, launch list, bench + reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/c5_all_tests.txt 2>&1
tail -5 gpurun_out/c5_all_tests.txt | cut -c1-300
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_case.py 2>&1 | grep -v "Host Frame\|Saved host backtrace\|^=========\s*$" | head -c 300000 > gpurun_out/c5_racecheck.txt
tail -6 gpurun_out/c5_racecheck.txt
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_case.py 2>&1 | tail -20 > gpurun_out/c5_memcheck.txt
tail -3 gpurun_out/c5_memcheck.txt
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'k2_fused|k2_update|k_strided|k_contig' -c 10 -o gpurun_out/c5_bench_full -f python bench.py --steps 1 --batch 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/c5_ncu_bench.log 2>&1
ls -la gpurun_out/c5_bench_full.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1200 --csv --log-file gpurun_out/c5_launches.csv python bench.py --steps 1 --batch 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/c5_launches.csv | head -12
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
tail -c 1200 gpurun_out/c5_bench.json; tail -3 gpurun_out/c5_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c5_bench_ref.json 2> gpurun_out/c5_bench_ref.err
cut -c1-900 gpurun_out/c5_bench_ref.json; tail -3 gpurun_out/c5_bench_ref.err
du -sh gpurun_out
