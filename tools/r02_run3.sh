#!/bin/bash
# round 2, GPU call 3 (2 GPUs): multi-GPU parity tests + 2-GPU bench line (C2, C4 sub-record, sharded search)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus2.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_multi_tests_2gpu.log 2>&1
tail -15 gpurun_out/r2_multi_tests_2gpu.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 4000 gpurun_out/r2_bench_n2.json; tail -20 gpurun_out/r2_bench_n2.err
