#!/bin/bash
# round 2, GPU call 17 (1 GPU): whole GPU suite after the list-order fix, breakdown of a 1 GiB mixed text,
# bench lines of the 256 MiB repetitive text and of the default workload (headline)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_gpu_tests_run17.log 2>&1
tail -8 gpurun_out/r2_gpu_tests_run17.log
timeout 600 python tools/ab2.py --workloads c4:1024,c1:64 --variants "" --steps 3 > gpurun_out/r2_ab_run17.txt 2>&1
cat gpurun_out/r2_ab_run17.txt
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --no-search > gpurun_out/r2_bench_c3_run17.json 2> gpurun_out/r2_bench_c3_run17.err
tail -c 600 gpurun_out/r2_bench_c3_run17.json; tail -3 gpurun_out/r2_bench_c3_run17.err
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_run17.json 2> gpurun_out/r2_bench_n1_run17.err
tail -c 1500 gpurun_out/r2_bench_n1_run17.json; tail -3 gpurun_out/r2_bench_n1_run17.err
