#!/bin/bash
# round 2, GPU call 5 (2 GPUs): full 2-GPU line with the fused exchange (C2 + C4 sub-record + sharded search), N=1 line
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29557 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_c.json 2> gpurun_out/r2_bench_n2_c.err
tail -c 6000 gpurun_out/r2_bench_n2_c.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n2_c.err | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-search --no-oracle-verify --cpu-sample-mib 16 > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err
tail -c 1200 gpurun_out/r2_bench_n1_c.json
