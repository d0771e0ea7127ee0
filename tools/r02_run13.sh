#!/bin/bash
# round 2, GPU call 13 (1 GPU): items-per-thread A/B around 18, full GPU suite + smoke at HEAD, final single-GPU line,
# refreshed ncu evidence for the 18-item radix pass
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_VARIANTS="_i16 _i17 _i19" AB_MIB=1024 timeout 900 bash tools/ab.sh > gpurun_out/r2_ab3.txt 2>&1
cat gpurun_out/r2_ab3.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_gpu_tests_1gpu_final.log 2>&1
tail -8 gpurun_out/r2_gpu_tests_1gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
tail -c 1500 gpurun_out/r2_bench_n1_final.json; tail -3 gpurun_out/r2_bench_n1_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-search --no-oracle-verify > gpurun_out/r2_ncu_launches_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"onesweep|pack_keys|init_ranks" -s 0 -c 9 -o gpurun_out/r2_onesweep18 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-search --no-oracle-verify > gpurun_out/r2_ncu18.log 2>&1
tail -2 gpurun_out/r2_ncu18.log
