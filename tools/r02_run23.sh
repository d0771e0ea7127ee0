#!/bin/bash
# round 2, GPU call 23 (1 GPU): launch list of the final headline command (kernel shares of the step)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-search --no-oracle-verify > gpurun_out/r2_ncu_launches_final2.log 2>&1
tail -2 gpurun_out/r2_ncu_launches_final2.log | cut -c1-300
