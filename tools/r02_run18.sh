#!/bin/bash
# round 2, GPU call 18 (1 GPU): A/B of the split key packing; launch list of a 1 GiB mixed text (where do 190 ms of
# rank work go?); ncu --set full of the in-group sort (256 MiB repetitive text), the search kernel and the packing
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/ab2.py --workloads c2:1024,c3:256 --variants "" --steps 5 > gpurun_out/r2_ab_run18_split.txt 2>&1
SAB_PACK_SPLIT=0 timeout 600 python tools/ab2.py --workloads c2:1024,c3:256 --variants "" --steps 5 > gpurun_out/r2_ab_run18_nosplit.txt 2>&1
cat gpurun_out/r2_ab_run18_split.txt gpurun_out/r2_ab_run18_nosplit.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_c4.csv python tools/ab2.py --workloads c4:1024 --variants "" --steps 1 --warmup 0 > gpurun_out/r2_ncu_launches_c4.log 2>&1
tail -3 gpurun_out/r2_ncu_launches_c4.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"group_sort_kernel" -s 6 -c 2 -o gpurun_out/r2_gsort python tools/ab2.py --workloads c3:256 --variants "" --steps 1 --warmup 0 > gpurun_out/r2_ncu_gsort.log 2>&1
tail -2 gpurun_out/r2_ncu_gsort.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"search_kernel|pack_keys|prefix_dir" -c 5 -o gpurun_out/r2_search python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-oracle-verify --patterns 10000000 > gpurun_out/r2_ncu_search.log 2>&1
tail -2 gpurun_out/r2_ncu_search.log
ls -la gpurun_out/*.ncu-rep
