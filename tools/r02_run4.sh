#!/bin/bash
# round 2, GPU call 4 (2 GPUs): fused peer-to-peer key exchange + single-sync planners: parity tests, A/B, full line
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_multi_tests_2gpu_b.log 2>&1
tail -15 gpurun_out/r2_multi_tests_2gpu_b.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
SAB_DIST_P2P=0 timeout 600 $TR --master-port 29556 bench.py --gpus 2 --steps 5 --warmup 3 --no-c4 --no-search --no-oracle-verify > gpurun_out/r2_bench_n2_nccl.json 2> gpurun_out/r2_bench_n2_nccl.err
tail -c 1800 gpurun_out/r2_bench_n2_nccl.json; tail -5 gpurun_out/r2_bench_n2_nccl.err
timeout 900 $TR --master-port 29557 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.err
tail -c 5000 gpurun_out/r2_bench_n2_b.json; tail -5 gpurun_out/r2_bench_n2_b.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-search --no-oracle-verify --cpu-sample-mib 16 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err
tail -c 2500 gpurun_out/r2_bench_n1_b.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-search --no-oracle-verify --no-cpu-baseline --api-gpus 2 > gpurun_out/r2_bench_api2.json 2> gpurun_out/r2_bench_api2.err
tail -c 1500 gpurun_out/r2_bench_api2.json; tail -3 gpurun_out/r2_bench_api2.err
