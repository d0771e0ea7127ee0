#!/bin/bash
# round 2, GPU call 19 (1 GPU): A/B after the restructured in-group sort and the cheaper split-filter marking
# (1 GiB DNA-like, 256 MiB repetitive, 1 GiB mixed), then the whole GPU suite, ncu --set full of the search kernel
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tools/ab2.py --workloads c2:1024,c3:256,c4:1024 --variants "" --steps 4 > gpurun_out/r2_ab_run19.txt 2>&1
cat gpurun_out/r2_ab_run19.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_gpu_tests_run19.log 2>&1
tail -8 gpurun_out/r2_gpu_tests_run19.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"search_kernel|prefix_dir" -c 3 -o gpurun_out/r2_search python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-oracle-verify > gpurun_out/r2_ncu_search.log 2>&1
tail -2 gpurun_out/r2_ncu_search.log
