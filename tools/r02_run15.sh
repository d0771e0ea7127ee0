#!/bin/bash
# round 2, GPU call 15 (1 GPU): the new in-group sort (warp sort of groups <= 512, owner tiles) -- direct test,
# parity subset, 256 MiB repetitive text; A/B of radix tile shapes and of a denser lazy-ISA directory on the 1 GiB text
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "native or golden or adversarial or group_sort or filter or shapes_16mib or random_mid" ) > gpurun_out/r2_gsort_tests.log 2>&1
tail -8 gpurun_out/r2_gsort_tests.log
timeout 600 python tools/ab2.py --workloads c3:256 --variants "" _mid256 --steps 4 > gpurun_out/r2_ab_c3.txt 2>&1
cat gpurun_out/r2_ab_c3.txt
timeout 900 python tools/ab2.py --workloads c2:1024 --variants "" _t512x9 _t384x12 _t512x10 _t512x8 _dir2 --steps 5 > gpurun_out/r2_ab_c2.txt 2>&1
cat gpurun_out/r2_ab_c2.txt
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --no-search > gpurun_out/r2_bench_c3_gsort.json 2> gpurun_out/r2_bench_c3_gsort.err
tail -c 1500 gpurun_out/r2_bench_c3_gsort.json; tail -3 gpurun_out/r2_bench_c3_gsort.err
