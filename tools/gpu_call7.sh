#!/bin/bash
# GPU call 7: final single-GPU evidence: all tests, smoke, k2check (fence / no fence), ncu captures, bench + reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/c7_all_tests.txt 2>&1
tail -4 gpurun_out/c7_all_tests.txt | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c7_smoke.txt 2>&1; tail -2 gpurun_out/c7_smoke.txt
( echo "== with the consumer fence (shipped)"; timeout 120 tools/k2check/k2_check 3932161 8 6 3 4; timeout 120 tools/k2check/k2_check 15728641 8 40 3 2
  echo "== without the fence (-DBK2_NO_WAR_FENCE)"; timeout 120 tools/k2check/k2_check_nofence 3932161 8 6 3 4; timeout 120 tools/k2check/k2_check_nofence 15728641 8 40 3 2 ) > gpurun_out/c7_k2check.txt 2>&1
grep -E "==|update rep|time" gpurun_out/c7_k2check.txt | head -30
timeout 600 ncu --set full --clock-control none -k regex:'k2_apply|k_sh_apply|k_cgl_apply|k_potrap|k_reduce|k_axpby|k_scale|k_gen|k_tail|k_lincomb' -c 44 -o gpurun_out/c7_tour -f python tools/kernel_tour.py > gpurun_out/c7_tour.log 2>&1
tail -2 gpurun_out/c7_tour.log; ls -la gpurun_out/c7_tour.ncu-rep
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'k2_fused|k2_update' -s 2400 -c 4 -o gpurun_out/c7_pair_late -f python bench.py --steps 3 --batch 10 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/c7_pair_late.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'k_strided|k_contig' -s 300 -c 3 -o gpurun_out/c7_fft_e8 -f python bench.py --steps 1 --batch 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/c7_fft_e8.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
tail -c 1500 gpurun_out/c7_bench.json; tail -3 gpurun_out/c7_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c7_bench_ref.json 2> gpurun_out/c7_bench_ref.err
cut -c1-400 gpurun_out/c7_bench_ref.json; tail -3 gpurun_out/c7_bench_ref.err
rm -f gpurun_out/c5_bench_full.ncu-rep
du -sh gpurun_out
