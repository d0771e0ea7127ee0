#!/bin/bash
# round 2, GPU call 14 (1 GPU): the whole GPU suite at HEAD (after the staged-copy fix) + smoke + final line
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_gpu_tests_1gpu_final.log 2>&1
tail -8 gpurun_out/r2_gpu_tests_1gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
tail -c 1200 gpurun_out/r2_bench_n1_final.json; tail -3 gpurun_out/r2_bench_n1_final.err
