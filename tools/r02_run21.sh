#!/bin/bash
# round 2, GPU call 21 (2 GPUs): the multi-GPU driver with the new in-group sort and key packing -- parity tests
# (default policy, complete inverse suffix array, one process driving both GPUs) and a verified 2-GPU bench line
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "29621 or 29623 or (single_process and 2)" ) > gpurun_out/r2_multi_tests_2gpu_run21.log 2>&1
tail -6 gpurun_out/r2_multi_tests_2gpu_run21.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29561 bench.py --gpus 2 --steps 3 --warmup 3 --no-c4 --no-search --no-oracle-verify > gpurun_out/r2_bench_n2_run21.json 2> gpurun_out/r2_bench_n2_run21.err
tail -c 2500 gpurun_out/r2_bench_n2_run21.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n2_run21.err | tail -5
