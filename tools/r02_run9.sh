#!/bin/bash
# round 2, GPU call 9 (2 GPUs): where is the cross-over between peer-to-peer and all-to-all rounds?
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for lim in 64000000 4000000; do
SAB_P2P_MAX_RECORDS=$lim timeout 600 $TR --master-port 29560 bench.py --gpus 2 --steps 5 --warmup 3 --no-c4 --no-search --no-oracle-verify > gpurun_out/r2_bench_n2_p2p$lim.json 2> gpurun_out/r2_bench_n2_p2p$lim.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n2_p2p$lim.json').read().strip().splitlines()[-1])
print('limit $lim', d['ms_per_step'], d['config']['phase_ms_rank0'])
PY
done
