#!/bin/bash
# round 2, GPU call 11 (1 GPU): the 256 MiB repetitive text (configs[2]): per-round picture, group-sort limit A/B, launch list
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "" _gs64 _gs128 _nofilter; do
  SAB200_LIB=$PWD/suffix_array_b200/libsab200$v.so timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-search --no-cpu-baseline --no-oracle-verify > gpurun_out/r2_c3$v.json 2> gpurun_out/r2_c3$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_c3$v.json').read().strip().splitlines()[-1])
print('variant "$v"', d['ms_per_step'], d['breakdown_ms'], d['group_sort'], d['verified'])
print('   active', d['config']['active']); print('   passes', d['config']['radix_passes'])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 0 --no-cpu-baseline --no-search --no-oracle-verify > gpurun_out/r2_ncu_c3.log 2>&1
tail -2 gpurun_out/r2_ncu_c3.log
