#!/bin/bash
# round 2, GPU call 20 (1 GPU): final single-GPU state -- A/B (static vs dynamic distribution of the warp sorts),
# whole GPU suite, smoke, headline bench line
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tools/ab2.py --workloads c2:1024,c3:256,c4:1024 --variants "" _dyn --steps 4 > gpurun_out/r2_ab_run20.txt 2>&1
cat gpurun_out/r2_ab_run20.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_gpu_tests_run20.log 2>&1
tail -8 gpurun_out/r2_gpu_tests_run20.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_run20.log 2>&1; tail -2 gpurun_out/r2_smoke_run20.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_run20.json 2> gpurun_out/r2_bench_n1_run20.err
tail -c 1200 gpurun_out/r2_bench_n1_run20.json; tail -3 gpurun_out/r2_bench_n1_run20.err
