#!/bin/bash
# round 2, GPU call 1: quick parity of the new radix pass, A/B of its variants on the 1 GiB text, ncu of the default
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "native or golden or adversarial or random_mid or baseline_shapes or filter_and_group" > gpurun_out/r2_quick_tests.log 2>&1
tail -3 gpurun_out/r2_quick_tests.log
AB_VARIANTS="_once _unf _i12 _i20 _once12" AB_MIB=1024 timeout 900 bash tools/ab.sh > gpurun_out/r2_ab1.txt 2>&1
cat gpurun_out/r2_ab1.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"onesweep" -s 3 -c 3 -o gpurun_out/r2_onesweep python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-search > gpurun_out/r2_ncu1.log 2>&1
tail -2 gpurun_out/r2_ncu1.log
