#!/bin/bash
# round 2, GPU call 7 (4 GPUs): multi-GPU parity tests with 4 ranks, fused bucket table on 4 GPUs, 4-GPU bench line
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -k "dist_construction or single_process or fused_buckets or sharded_over_replicas" > gpurun_out/r2_multi_tests_4gpu.log 2>&1
tail -12 gpurun_out/r2_multi_tests_4gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29558 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
tail -c 6000 gpurun_out/r2_bench_n4.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n4.err | tail -8
