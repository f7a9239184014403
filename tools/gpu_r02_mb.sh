#!/bin/bash
# launch bounds of pass 2 of the blocked insert: minimum resident CTAs per SM 1 (shipped: 55 registers, 4 CTAs),
# 5 and 6 (dev builds of kind 1)
mkdir -p gpurun_out
for lib in libcuco_b200.so libcuco_b200_dev_mb5.so libcuco_b200_dev_mb6.so; do
  CUCO_B200_LIB=$PWD/cucollections_b200/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-points --no-cpu-baseline \
    > gpurun_out/r02_minblocks_$lib.json 2> gpurun_out/r02_minblocks_$lib.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02_minblocks_$lib.json'))
    print('$lib', {k: round(d[k], 3) for k in ('value', 'insert_gops', 'insert_ms', 'insert_ms_best', 'find_ms')})
except Exception as e:
    print('$lib failed', e)
PY
done
