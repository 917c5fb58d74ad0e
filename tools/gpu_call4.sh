#!/bin/bash
# GPU call 4: warm per-kernel times of the transform kernels for every E, all GPU tests, racecheck, ncu full captures, bench
mkdir -p gpurun_out
for le in 2 3 4 5; do
  BK_FFT_LOGE=$le timeout 300 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -k regex:'k_strided|k_contig' -s 30 -c 30 --csv --log-file gpurun_out/c4_fft_warm_e$le.csv python tools/bench_precond.py 1024 > /dev/null 2>&1
  BK_FFT_LOGE=$le timeout 300 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -k regex:'k_strided|k_contig' -s 30 -c 30 --csv --log-file gpurun_out/c4_fft512_warm_e$le.csv python tools/bench_precond.py 512 > /dev/null 2>&1
done
python tools/launch_summary.py gpurun_out/c4_fft_warm_e2.csv; python tools/launch_summary.py gpurun_out/c4_fft_warm_e3.csv; python tools/launch_summary.py gpurun_out/c4_fft_warm_e4.csv; python tools/launch_summary.py gpurun_out/c4_fft_warm_e5.csv
python tools/launch_summary.py gpurun_out/c4_fft512_warm_e2.csv; python tools/launch_summary.py gpurun_out/c4_fft512_warm_e3.csv; python tools/launch_summary.py gpurun_out/c4_fft512_warm_e5.csv
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/c4_all_tests.txt 2>&1
tail -30 gpurun_out/c4_all_tests.txt | cut -c1-300
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_case.py > gpurun_out/c4_racecheck.txt 2>&1
tail -12 gpurun_out/c4_racecheck.txt
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/c4_memcheck.txt 2>&1
tail -4 gpurun_out/c4_memcheck.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k2_fused|k2_update|k2_apply|k_strided|k_contig|k_reduce|k_lincomb|k_axpby' -c 40 -o gpurun_out/c4_bench_full -f python bench.py --steps 1 --batch 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/c4_ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/c4_launches.csv python bench.py --steps 1 --batch 3 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/c4_launches.csv | head -20
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
tail -c 1500 gpurun_out/c4_bench.json; tail -3 gpurun_out/c4_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c4_bench_ref.json 2> gpurun_out/c4_bench_ref.err
cat gpurun_out/c4_bench_ref.json | cut -c1-1500; tail -3 gpurun_out/c4_bench_ref.err
