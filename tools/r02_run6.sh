#!/bin/bash
# round 2, GPU call 6 (1 GPU): the whole GPU test suite at HEAD + the single-GPU line
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_gpu_tests_1gpu.log 2>&1
tail -8 gpurun_out/r2_gpu_tests_1gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-oracle-verify --cpu-sample-mib 32 > gpurun_out/r2_bench_n1_d.json 2> gpurun_out/r2_bench_n1_d.err
tail -c 3500 gpurun_out/r2_bench_n1_d.json; tail -3 gpurun_out/r2_bench_n1_d.err
