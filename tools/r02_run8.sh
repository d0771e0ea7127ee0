#!/bin/bash
# round 2, GPU call 8 (2 GPUs): peer-to-peer rounds: parity tests (all exchange forms), 2-GPU line
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_multi_tests_2gpu_c.log 2>&1
tail -15 gpurun_out/r2_multi_tests_2gpu_c.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29559 bench.py --gpus 2 --steps 5 --warmup 3 --no-c4 --no-search --no-oracle-verify > gpurun_out/r2_bench_n2_d.json 2> gpurun_out/r2_bench_n2_d.err
tail -c 2500 gpurun_out/r2_bench_n2_d.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n2_d.err | tail -8
