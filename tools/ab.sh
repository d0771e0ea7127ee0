#!/bin/bash
# developer tool: A/B the variant builds of libsab200 on one box (bench at 256 MiB, device-resident)
for v in "" ${AB_VARIANTS:-_lb1 _lb4 _ms _v5}; do
  lib=suffix_array_b200/libsab200$v.so
  [ -f $lib ] || continue
  echo -n "variant '$v': "
  SAB200_LIB=$PWD/$lib python bench.py --n-mib ${AB_MIB:-256} --steps 2 --warmup 2 --no-search --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'],'MB/s', d['ms_per_step'],'ms', 'pass GB/s',d['roofline']['achieved'], d['breakdown_ms'])"
done
