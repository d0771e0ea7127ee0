#!/bin/bash
# round 2, GPU call 10 (8 GPUs): the 8-GPU line (C2 + C4 sub-record + search on 8 replicas), threshold A/B,
# sab200_saca(ngpus = 8) in one process
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -c 6000 gpurun_out/r2_bench_n8.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n8.err | tail -8
SAB_P2P_MAX_RECORDS=2000000 timeout 600 $TR --master-port 29562 bench.py --gpus 8 --steps 5 --warmup 3 --no-c4 --no-search --no-oracle-verify > gpurun_out/r2_bench_n8_p2p2m.json 2> gpurun_out/r2_bench_n8_p2p2m.err
tail -c 2000 gpurun_out/r2_bench_n8_p2p2m.json
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "single_process" > gpurun_out/r2_multi_tests_8gpu.log 2>&1
tail -5 gpurun_out/r2_multi_tests_8gpu.log
