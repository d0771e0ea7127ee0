#!/bin/bash
# round 2, GPU call 12 (1 GPU): ncu --set full of the non-radix kernels of the step; A/B of a few more radix-pass shapes
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pack_keys|init_ranks|gather_rank2|rerank|radix_hist|group_sort" -s 6 -c 8 -o gpurun_out/r2_others python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-search --no-oracle-verify > gpurun_out/r2_ncu_others.log 2>&1
tail -2 gpurun_out/r2_ncu_others.log
AB_VARIANTS="_hwm4 _i18 _i14 _lb2" AB_MIB=1024 timeout 900 bash tools/ab.sh > gpurun_out/r2_ab2.txt 2>&1
cat gpurun_out/r2_ab2.txt
