#!/bin/bash
# round 2, GPU call 16 (1 GPU): whole GPU suite at HEAD (tuned in-group sort, prefix directory of the resident index,
# pipelined batched search), A/B of the in-group sort on the 1 GiB / 256 MiB texts, search with / without the
# directory and with 4 lanes per pattern
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_gpu_tests_run16.log 2>&1
tail -8 gpurun_out/r2_gpu_tests_run16.log
timeout 600 python tools/ab2.py --workloads c3:256,c2:1024 --variants "" --steps 5 > gpurun_out/r2_ab_run16.txt 2>&1
cat gpurun_out/r2_ab_run16.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-oracle-verify > gpurun_out/r2_bench_run16.json 2> gpurun_out/r2_bench_run16.err
tail -c 3000 gpurun_out/r2_bench_run16.json; tail -3 gpurun_out/r2_bench_run16.err
SAB_SEARCH_DIR=0 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-oracle-verify > gpurun_out/r2_bench_run16_nodir.json 2> gpurun_out/r2_bench_run16_nodir.err
tail -3 gpurun_out/r2_bench_run16_nodir.err
SAB200_LIB=$PWD/suffix_array_b200/libsab200_g4.so timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-oracle-verify > gpurun_out/r2_bench_run16_g4.json 2> gpurun_out/r2_bench_run16_g4.err
tail -3 gpurun_out/r2_bench_run16_g4.err
python - <<'PY'
import json
for f in ("run16", "run16_nodir", "run16_g4"):
    try:
        d = json.load(open("gpurun_out/r2_bench_%s.json" % f))
        s = d["search"]
        print(f, d["ms_per_step"], {k: (s[k]["queries_per_s_kernel"], s[k]["queries_per_s_host_abi"], s[k]["roofline"]["probes_per_query"], s[k]["oracle_equal_on_sample"]) for k in ("alphabet_hybrid", "raw_byte_hybrid")}, s.get("prefix_directory"), s.get("index_create_ms"))
    except Exception as e:
        print(f, "failed", e)
PY
