#!/bin/bash
# round 2, GPU call 22 (1 GPU): BASELINE configs[3] (3.9 GiB mixed text) on ONE GPU with the final build -- the base of the
# c4 speed-ups on the multi-GPU bench lines
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 280 python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/r2_c4_1gpu_final.json 2> gpurun_out/r2_c4_1gpu_final.err
tail -c 1500 gpurun_out/r2_c4_1gpu_final.json; tail -3 gpurun_out/r2_c4_1gpu_final.err
