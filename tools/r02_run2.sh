#!/bin/bash
# round 2, GPU call 2 (1 GPU): new bench line (verification, pageable e2e, C5 as specified), C4 on one GPU,
# reference arm with one full-size construction, launch list
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nproc > gpurun_out/r2_nproc.txt; free -g >> gpurun_out/r2_nproc.txt; df -h /dev/shm >> gpurun_out/r2_nproc.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -c 3000 gpurun_out/r2_bench_n1.json; tail -5 gpurun_out/r2_bench_n1.err
timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/r2_c4_n1.json 2> gpurun_out/r2_c4_n1.err
cat gpurun_out/r2_c4_n1.json; tail -5 gpurun_out/r2_c4_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-search --no-oracle-verify > gpurun_out/r2_ncu_launches.log 2>&1
tail -2 gpurun_out/r2_ncu_launches.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 --ref-full > gpurun_out/r2_ref_full.json 2> gpurun_out/r2_ref_full.err
cat gpurun_out/r2_ref_full.json
